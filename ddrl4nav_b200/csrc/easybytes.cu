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

#include <algorithm>
#include <cstring>

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

// ---- reply encode (encode_forward_return_data, easybytes.py:77-109) on the device -----------------------------------------
// Reply of env process j = blocks [actions[j*nb:(j+1)*nb], logps[...], values[:, j*nb:(j+1)*nb]] with the data-block headers
// of encode_data (type >h = 3 (f32) | count >I | ndim >I | shape >I x ndim).  Every reply has the same layout, so the kernel
// writes headers and payload of all replies into one byte buffer (one D2H copy, the host only cuts it per env process).
struct ReplyLayout {
  int nb, A, V, B;                  // rows per env process, action columns (0: 1-D actions), value rows, batch rows
  int off_a, off_l, off_v, total;   // byte offsets of the three PAYLOADS inside one reply, reply size
  unsigned char hdr[3][24];         // headers of the three blocks (hdr_len bytes each valid)
  int hdr_len[3];
};

__global__ void __launch_bounds__(256) easybytes_encode_replies_kernel(const float* __restrict__ actions, const float* __restrict__ logps,
                                                                       const float* __restrict__ values, ReplyLayout L, int n_env,
                                                                       uint8_t* __restrict__ out) {
  const int ac = L.A > 0 ? L.A : 1;
  const int fa = L.nb * ac, fl = L.nb, fv = L.V * L.nb;          // floats per reply and block
  const int per = fa + fl + fv;
  const int hb = L.hdr_len[0] + L.hdr_len[1] + L.hdr_len[2];
  const long long work = (long long)n_env * (per + hb);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < work; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / (per + hb));
    int r = (int)(t - (long long)j * (per + hb));
    uint8_t* o = out + (long long)j * L.total;
    if (r < hb) {                                                 // one header byte
      int k = 0;
      if (r >= L.hdr_len[0]) { r -= L.hdr_len[0]; k = 1; if (r >= L.hdr_len[1]) { r -= L.hdr_len[1]; k = 2; } }
      const int pay = k == 0 ? L.off_a : (k == 1 ? L.off_l : L.off_v);
      o[pay - L.hdr_len[k] + r] = L.hdr[k][r];
      continue;
    }
    r -= hb;
    float v;
    int dst;
    if (r < fa) { v = actions[(long long)j * fa + r]; dst = L.off_a + 4 * r; }
    else if (r < fa + fl) { r -= fa; v = logps[(long long)j * L.nb + r]; dst = L.off_l + 4 * r; }
    else { r -= fa + fl; const int vr = r / L.nb, i = r - vr * L.nb; v = values[(long long)vr * L.B + (long long)j * L.nb + i]; dst = L.off_v + 4 * r; }
    const unsigned u = __float_as_uint(v);                        // little-endian payload, only byte-aligned on the wire
    o[dst] = (uint8_t)u; o[dst + 1] = (uint8_t)(u >> 8); o[dst + 2] = (uint8_t)(u >> 16); o[dst + 3] = (uint8_t)(u >> 24);
  }
}

static int put_header(unsigned char* h, int count, int ndim, const int* shape) {
  auto be32 = [](unsigned char* p, unsigned v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
  h[0] = 0; h[1] = 3;                                              // >h 3 = float32 (easybytes.py:21-26)
  be32(h + 2, (unsigned)count); be32(h + 6, (unsigned)ndim);
  for (int i = 0; i < ndim; ++i) be32(h + 10 + 4 * i, (unsigned)shape[i]);
  return 10 + 4 * ndim;
}

}  // namespace ddrl

using namespace ddrl;

extern "C" int64_t ddrl_easybytes_reply_bytes(int nb, int act_cols, int V) {
  if (nb < 0 || act_cols < 0 || V < 1) return DDRL_E_ARG;
  const int ac = act_cols > 0 ? act_cols : 1;
  return (int64_t)(10 + 4 * (act_cols > 0 ? 2 : 1)) + 4LL * nb * ac + 14 + 4LL * nb + 22 + 4LL * V * nb;
}

extern "C" int ddrl_easybytes_encode_replies(const float* actions, int act_cols, const float* logps, const float* values, int V,
                                             int B, int n_env, int nb, uint8_t* out, void* stream) {
  if (n_env < 0 || nb < 0 || V < 1 || act_cols < 0 || (long long)n_env * nb > B) return DDRL_E_ARG;
  if (n_env == 0) return DDRL_OK;
  if (!actions || !logps || !values || !out) return DDRL_E_ARG;
  ReplyLayout L;
  memset(&L, 0, sizeof(L));
  L.nb = nb; L.A = act_cols; L.V = V; L.B = B;
  const int ac = act_cols > 0 ? act_cols : 1;
  const int sa[2] = {nb, act_cols}, sl[1] = {nb}, sv[3] = {V, nb, 1};
  L.hdr_len[0] = put_header(L.hdr[0], nb * ac, act_cols > 0 ? 2 : 1, sa);
  L.hdr_len[1] = put_header(L.hdr[1], nb, 1, sl);
  L.hdr_len[2] = put_header(L.hdr[2], V * nb, 3, sv);
  L.off_a = L.hdr_len[0];
  L.off_l = L.off_a + 4 * nb * ac + L.hdr_len[1];
  L.off_v = L.off_l + 4 * nb + L.hdr_len[2];
  L.total = L.off_v + 4 * V * nb;
  const long long work = (long long)n_env * (nb * (ac + 1 + V) + L.hdr_len[0] + L.hdr_len[1] + L.hdr_len[2]);
  const int blocks = (int)std::min<long long>((work + 255) / 256, 4LL * kNumSMs);
  easybytes_encode_replies_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(actions, logps, values, L, n_env, out);
  DDRL_LAUNCHED("easybytes_encode_replies_kernel");
  return DDRL_OK;
}

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
