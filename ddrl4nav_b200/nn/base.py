"""`Basenn` / `PreNet` with the reference's interface (USTC_lab/nn/base.py:23-153).

Weight (de)serialisation keeps the reference's wire format (base.py:38-54): for every
``named_parameters()`` entry, in order, ``>I ndim | >I*ndim shape | raw fp32 bytes``.  Unlike
the reference (one ``.cpu()`` and one ``bytes +=`` per parameter) the blob is produced from a
single device->host copy of the flat parameter buffer.
"""
import struct

import numpy as np
import torch
import torch.nn


class Basenn(torch.nn.Module):
    def __init__(self, config, config_nn):
        super().__init__()
        self.conn = self._connect_redis(getattr(config, "MIDDLE_REDIS_HOST", None), getattr(config, "MIDDLE_REDIS_PORT", None))
        self.pipe = self.conn.pipeline() if self.conn is not None else None
        self.model_key = getattr(config, "TASK_NAME", "") + getattr(config, "MODULE_KEY", "MODEL")
        self.device = getattr(config, "DEVICE", "cuda")
        self.model_dtype = getattr(config_nn, "MODULE_NUMPY_DTYPE", np.float32)
        self.model_dtype_bytes = getattr(config_nn, "MODULE_BITS", 32) // 8
        self.model_tensor_dtype = getattr(config_nn, "MODULE_TENSOR_DTYPE", torch.float32)

    def _connect_redis(self, host, port):
        try:
            import redis
        except ImportError:            # no redis client installed: weight sync needs an injected connection
            return None
        return redis.Redis(host=host, port=port)

    # -- wire format (base.py:38-54) --------------------------------------------------------
    def _encode_wb(self, wb_np: np.ndarray) -> bytes:
        shape = wb_np.shape
        return struct.pack(">I", len(shape)) + struct.pack(">%dI" % len(shape), *shape) + wb_np.tobytes()

    def _decode_wb(self, wb_bytes):
        ndim = struct.unpack_from(">I", wb_bytes, 0)[0]
        shape = struct.unpack_from(">%dI" % ndim, wb_bytes, 4)
        count = int(np.prod(shape)) if ndim else 1
        head = 4 + 4 * ndim
        wb = np.frombuffer(wb_bytes, dtype=self.model_dtype, offset=head, count=count).reshape(shape)
        return wb, head + count * self.model_dtype_bytes

    def _params_to_host(self):
        """[(name, np.ndarray)] in named_parameters() order; subclasses with a flat buffer override."""
        return [(k, v.detach().cpu().numpy()) for k, v in self.named_parameters()]

    def model_bytes(self) -> bytes:
        return b"".join(self._encode_wb(a) for _, a in self._params_to_host())

    def nn2redis(self, pipe, update_key, key=None):
        pipe.set(key if key else self.model_key, self.model_bytes())
        pipe.incr(update_key)
        pipe.execute()

    def load_model_bytes(self, model_bytes):
        view = memoryview(model_bytes)
        index = 0
        with torch.no_grad():
            for _, p in self.named_parameters():
                wb, used = self._decode_wb(view[index:])
                index += used
                p.copy_(torch.from_numpy(np.ascontiguousarray(wb)).to(p.device))
        self._weights_changed()

    def updatenn_by_redis(self, conn, key=None):
        self.load_model_bytes(conn.get(key if key else self.model_key))

    def updatenn_by_file(self, file_path: str):
        self.load_state_dict(torch.load(file_path))
        self._weights_changed()

    def updatenn(self, path: str, conn=None):
        if path.startswith("redis"):
            assert conn is not None
            self.updatenn_by_redis(conn, path.split("://")[-1])
        elif path.startswith("file"):
            self.updatenn_by_file(path.split("://")[-1])

    def _weights_changed(self):
        pass

    def states_normalization(self, states):
        pass

    def imitation_learning(self, *args, **kwargs):
        raise NotImplementedError("imitation pre-training is outside the B200 hot path (SURVEY 2 #1: keep the "
                                  "reference's nn/base.py:109-150 for it)")


class PreNet(torch.nn.Module):
    """Encoder base.  Encoders own parameters (names/shapes/init = reference) and a static
    description of their family; the arithmetic runs inside PPO's fused engine."""
    ARCH = None

    def __init__(self):
        super().__init__()

    def engine_in_ch(self) -> int:
        raise NotImplementedError

    def forward(self, states):
        from .._lib import DDRLError
        raise DDRLError("%s runs inside ddrl4nav_b200.nn.PPO's fused CUDA engine (PPO.forward / PPO.learn); "
                        "a stand-alone eager forward does not exist (no PyTorch fallback)." % type(self).__name__)
