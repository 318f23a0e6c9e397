// EasyBytes payload -> fp32 state tensors, on the device.
//
// Replaces the decode half of USTC_lab/data/easybytes.py on the Forward path: decode_forward_states (:114-139) slices
// the Redis payload per env process with np.frombuffer, np.concatenate's every state slot over the processes, and the
// Forward thread then converts each slot to fp32 (server/forward.py:128-131).  Here the host only reads the headers
// (a few dozen bytes per message, ddrl4nav_b200/data/easybytes.py); the raw payload goes to the GPU in ONE copy and one
// kernel does slice + concatenate + dtype conversion: segment i = `count` elements of wire type `dtype`
// (1 u8, 2 f16, 3 f32, 4 f64; little-endian, easybytes.py:21-26) at byte `src_off`, written as fp32 at `dst_off`.
// Pure byte/convert work, HBM-bound: (elem size + 4) bytes per element.  Wire data is only byte-aligned
// (headers are 10 + 4*ndim bytes), so elements are assembled from bytes unless the segment happens to be aligned.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ddrl {

struct Seg {
  unsigned long long src_off, dst_off;
  unsigned int count, dtype;
};

__device__ __forceinline__ float load_elem(const uint8_t* p, unsigned dtype, bool aligned) {
  switch (dtype) {
    case 1: return (float)p[0];
    case 2: {
      unsigned short h = aligned ? *reinterpret_cast<const unsigned short*>(p) : (unsigned short)(p[0] | (p[1] << 8));
      return __half2float(__ushort_as_half(h));
    }
    case 3: {
      unsigned int u = aligned ? *reinterpret_cast<const unsigned int*>(p)
                               : (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
      return __uint_as_float(u);
    }
    default: {
      unsigned long long u;
      if (aligned) u = *reinterpret_cast<const unsigned long long*>(p);
      else {
        u = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) u |= (unsigned long long)p[b] << (8 * b);
      }
      return (float)__longlong_as_double((long long)u);       // cvt.rn.f32.f64 = numpy's astype(float32)
    }
  }
}

// grid.y = segment (grid-strided), grid.x * block = elements of the segment (grid-strided)
__global__ void __launch_bounds__(256) easybytes_decode_kernel(const uint8_t* __restrict__ payload, const Seg* __restrict__ segs,
                                                               int nseg, float* __restrict__ dst) {
  for (int sg = blockIdx.y; sg < nseg; sg += gridDim.y) {
    const Seg s = segs[sg];
    const unsigned esz = s.dtype == 1 ? 1u : (s.dtype == 2 ? 2u : (s.dtype == 3 ? 4u : 8u));
    const uint8_t* src = payload + s.src_off;
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & (esz - 1)) == 0;
    float* out = dst + s.dst_off;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < s.count; i += gridDim.x * blockDim.x)
      out[i] = load_elem(src + (size_t)i * esz, s.dtype, aligned);
  }
}

}  // namespace ddrl

using namespace ddrl;

extern "C" int ddrl_easybytes_decode(const uint8_t* payload, const void* segs, int nseg, unsigned int max_count, float* dst,
                                     void* stream) {
  if (nseg < 0 || (nseg > 0 && (!payload || !segs || !dst))) return DDRL_E_ARG;
  if (nseg == 0 || max_count == 0) return DDRL_OK;
  static_assert(sizeof(Seg) == 24, "segment record layout is part of the ABI");
  const unsigned bx = (unsigned)std::min<long long>(((long long)max_count + 255) / 256, 4LL * kNumSMs);
  const unsigned by = (unsigned)std::min(nseg, 65535);
  easybytes_decode_kernel<<<dim3(bx, by), 256, 0, (cudaStream_t)stream>>>(payload, reinterpret_cast<const Seg*>(segs), nseg, dst);
  DDRL_LAUNCHED("easybytes_decode_kernel");
  return DDRL_OK;
}
