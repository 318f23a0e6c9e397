"""Thread-level drop-ins for the two Redis-fed modules (SURVEY 8b "Modules"):

    ForwardThread(net, predictor_id, logger, easy_bytes, exit_flag, configs)      server/forward.py:20-29,107-183
    BackwardGetDataThread / BackwardTrainThread(net, training_data_queue, trainer_id, logger, easy_bytes,
                                                exit_flag, configs)              server/backward.py:68-76,141-217
    BackwardQueue                                                                 server/backward.py:42-65

Same constructor contracts, same Redis keys / ordering of Redis operations, same logger calls as the reference threads;
what runs between the Redis pop and the Redis push is the device path of this package:

  * Forward tick: the popped payload goes through ``ForwardModule.step_bytes_replies`` (header parse on the host, ONE H2D
    copy in wire dtypes -- streamed in chunks of whole messages when large --, decode + encoders + heads + sampling +
    log-prob + value + reply encode on the device, ONE D2H copy of the reply bytes).  ``encode_forward_replies`` is the
    host-side equivalent of the reply encoder (``encode_forward_return_data``), kept for arrays that are already on the host.
  * Trainer: ``BackwardGetDataThread`` queues the RAW payload (decoding it on the host is exactly the work the device
    decoder removes); ``BackwardQueue.get`` hands the payloads of one batch to ``DeviceEasyBytes.decode_backward_batch``
    (one staged copy + one kernel = decode + ``Experience.batch_data``), so ``train_data.to_tensor`` has nothing left to do.

``redis`` is imported when a thread is constructed (the reference imports it at module import); pass
``configs["redis_factory"]`` (a callable ``(host, port) -> connection``) to inject a connection, e.g. in tests.
The loops end when ``exit_flag.value != b'0'`` or when a blocking pop times out with nothing (the reference would raise
a TypeError on ``None[1]`` there)."""
import queue as _queue
import struct
import time
from collections import defaultdict
from threading import Thread
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from ..data.easybytes import DeviceEasyBytes, TYPE_SIZE
from ..data.experience import Experience
from .forward import ForwardModule

_NP_CODE = {np.dtype(np.uint8): 1, np.dtype(np.float16): 2, np.dtype(np.float32): 3, np.dtype(np.float64): 4}


def _redis_factory(configs):
    f = configs.get("redis_factory") if isinstance(configs, dict) else None
    if f is not None:
        return f
    import redis                                   # same dependency as the reference threads
    return lambda host, port: redis.Redis(host, port)


def _block_header(dtype, shape) -> bytes:
    """easybytes.py:62-75: type (>h), count (>I), ndim (>I), shape (>I x ndim)."""
    count = int(np.prod(shape, dtype=np.int64))
    return struct.pack(">hII", _NP_CODE[np.dtype(dtype)], count, len(shape)) + struct.pack(">" + "I" * len(shape), *shape)


def encode_forward_replies(arrays: Sequence[np.ndarray], env_batch_nums: Sequence[int]) -> List[bytes]:
    """``EasyBytes.encode_forward_return_data`` (easybytes.py:77-109): per env process the blocks
    [actions[i0:i1], logps[i0:i1], values[:, i0:i1], extras[i0:i1]...]; byte-identical output."""
    out, i0 = [], 0
    for nb in env_batch_nums:
        i1 = i0 + nb
        parts = []
        for k, a in enumerate(arrays):
            sl = a[:, i0:i1] if k == 2 else a[i0:i1]
            parts.append(_block_header(sl.dtype, sl.shape))
            parts.append(np.ascontiguousarray(sl).tobytes())
        out.append(b"".join(parts))
        i0 = i1
    return out


class ForwardThread(Thread):
    def __init__(self, net, predictor_id: int, logger, easy_bytes, exit_flag, configs):
        super().__init__()
        self.daemon = True
        self.net = net
        config, config_nn, config_env = configs["config"], configs["config_nn"], configs["config_env"]
        self.predictor_id = predictor_id
        self.easy_bytes = easy_bytes
        self.logger_f = logger
        self.logger_f.update_tensor_tags("predict/", predictor_id)
        connect = _redis_factory(configs)
        self.conn_pre = connect(config.PREDICTOR_REDIS_HOST, config.PREDICTOR_REDIS_PORT)
        self.pipe_pre = self.conn_pre.pipeline()
        self.conn_middle = connect(config.MIDDLE_REDIS_HOST, config.MIDDLE_REDIS_PORT)
        self.pipe_middle = self.conn_middle.pipeline()
        self.exit_flag = exit_flag
        self.tensortype = config_nn.MODULE_TENSOR_DTYPE
        self.nptype = config_nn.MODULE_NUMPY_DTYPE
        self.device = config_nn.DEVICE
        self.forward_states_key = config.TASK_NAME + config.PREDICTING_STATES_KEY
        self.pre_actionkey = config.TASK_NAME + config.PRE_ACTIONS_KEY
        self.update_tag = config.TASK_NAME + config.UPDATE_TAG_KEY
        self.env_dict_key = config.TASK_NAME + config.ENV_NUM_DICT_KEY
        self.env_dict = {str(k): int(v) for k, v in self.conn_pre.hgetall(self.env_dict_key).items()}
        self.episode = 0
        self.timeout = config.TIME_OUT
        self.action_dim = config_nn.ACTIONS_DIM
        self.play_mode = config.PLAY_MODE or config.DEMONSTRATE_MODE
        self.config, self.config_nn = config, config_nn
        self.sync = config.SYNC
        self.train_lock_key = config.TASK_NAME + config.TRAIN_LOCK_KEY
        self.batch_num_per_env = config_env["batch_num_per_env"]
        self.agent_num_per_env = config_env["agent_num_per_env"]
        if getattr(net, "rnd", None):
            raise ValueError("RND / GAIL reply columns are outside the B200 hot path: use the reference ForwardThread")
        self.module = ForwardModule(net, play_mode=self.play_mode, nptype=self.nptype, device=self.device)
        self._pre_update = 0

    def check_demonstrate(self) -> bool:
        if self.play_mode:
            self.net.updatenn(path=self.config.DEMONSTRATE_LOAD_PATH, conn=self.conn_middle)
            return True
        return False

    def state2tensor(self, states):
        """Kept for callers of the reference API; the tick itself never materialises host arrays."""
        for i in range(len(states)):
            states[i] = torch.as_tensor(np.asarray(states[i])).to(device=self.device, dtype=self.tensortype)

    def tick(self, byte_states, draw=None) -> int:
        """Everything ``run`` does for one popped payload (forward.py:118-181); returns the number of replies pushed."""
        while self.sync and int(self.conn_middle.get(self.train_lock_key)) == 1:
            time.sleep(0.1)
        t0 = time.time()
        if not self.play_mode:
            train_update = int(self.conn_middle.get(self.update_tag))
            if train_update > self._pre_update:
                self.net.updatenn_by_redis(self.conn_middle)
                self._pre_update = train_update
        per_env = self.batch_num_per_env * self.agent_num_per_env
        # decode + encoders + heads + sampling + reply encode on the device; one H2D of the wire bytes, one D2H of the replies
        env_ids, replies = self.module.step_bytes_replies(byte_states, per_env, draw=draw)
        for env_id, payload in zip(env_ids, replies):
            self.pipe_pre.lpush(self.pre_actionkey.format(env_id), payload)
        self.pipe_pre.execute()
        self.logger_f.add(((time.time() - t0) * 1000, self.episode), "ForwardTime-ms")
        self.episode += 1
        return len(replies)

    def run(self):
        if not self.check_demonstrate():
            while not self.conn_middle.get(self.update_tag):                 # wait for the trainer's first weights
                if self.exit_flag.value != b"0":
                    return
                time.sleep(0.3)
        while self.exit_flag.value == b"0":
            item = self.conn_pre.blpop(self.forward_states_key, timeout=self.timeout * 10)
            if item is None:
                break
            self.tick(item[1])
        print("forward exit !", flush=True)


def _batch_logger(dicts: List[Dict]) -> Dict:
    return {k: float(np.mean([d[k] for d in dicts])) for k in dicts[0]} if dicts else {}


def _payload_rows(payload) -> int:
    """Rows of a training payload = leading extent of its first state block (header only, no decode)."""
    ndim = struct.unpack_from(">I", payload, 8 + 6)[0]
    if ndim < 1:
        raise ValueError("EasyBytes: training payload whose first state block is a scalar")
    return struct.unpack_from(">I", payload, 8 + 10)[0]


class BackwardQueue:
    """``get(batch_size)`` pops payloads until they hold `batch_size` rows and returns (Experience of fp32 DEVICE tensors,
    averaged logger dict) -- BackwardQueue.get + Experience.batch_data + to_tensor of the reference in one device pass.
    Already-decoded ``(Experience, dict)`` items (the reference producer) are accepted as well."""

    def __init__(self, device="cuda"):
        self.q = _queue.Queue()
        self.device = device
        self._eb = None

    def put(self, data, *args) -> None:
        self.q.put(data, *args)

    def get(self, batch_size, *args) -> Tuple[Experience, Dict]:
        rows, payloads, exps, dicts = 0, [], [], []
        while rows < batch_size:
            item = self.q.get(*args)
            if isinstance(item, (bytes, bytearray, memoryview)):
                payloads.append(item)
                rows += _payload_rows(item)
            else:
                exp, d = item
                exps.append(exp)
                rows += len(exp)
                if len(d):
                    dicts.append(d)
        if payloads:
            if self._eb is None:
                self._eb = DeviceEasyBytes(self.device)
            exp, loggers = self._eb.decode_backward_batch(payloads)
            dicts.extend(d for d in loggers if len(d))
            if not exps:
                return exp, _batch_logger(dicts)
            exp_host = Experience.batch_data(exps)
            exp_host.to_tensor(device=self.device)
            cat = lambda a, b, ax=0: torch.cat([a, b], dim=ax)
            exp = Experience(states=[cat(a, b) for a, b in zip(exp.states, exp_host.states)], advs=cat(exp.advs, exp_host.advs),
                             actions=cat(exp.actions, exp_host.actions), old_logps=cat(exp.old_logps, exp_host.old_logps),
                             values=cat(exp.values, exp_host.values, 1))
            return exp, _batch_logger(dicts)
        return Experience.batch_data(exps), _batch_logger(dicts)


class BackwardThread(Thread):
    def __init__(self, net, training_data_queue, trainer_id: int, logger, easy_bytes, exit_flag, configs):
        super().__init__()
        config, config_nn = configs["config"], configs["config_nn"]
        self.net = net
        self.training_data_queue = training_data_queue
        self.easy_bytes = easy_bytes
        connect = _redis_factory(configs)
        self.conn_train = connect(config.TRAINER_REDIS_HOST, config.TRAINER_REDIS_PORT)
        self.pipe_train = self.conn_train.pipeline()
        self.conn_middle = connect(config.MIDDLE_REDIS_HOST, config.MIDDLE_REDIS_PORT)
        self.pipe_middle = self.conn_middle.pipeline()
        self.data_key = config.TASK_NAME + config.TRAINING_DATA_KEY
        self.logger_f = logger
        self.logger_f.update_tensor_tags("train/", trainer_id)
        self.tensortype = config_nn.MODULE_TENSOR_DTYPE
        self.nptype = config_nn.MODULE_NUMPY_DTYPE
        self.device = config_nn.DEVICE
        self.min_batch_size = config_nn.TRAINING_MIN_BATCH
        self.update_tag = config.TASK_NAME + config.UPDATE_TAG_KEY
        self.train_lock_key = config.TASK_NAME + config.TRAIN_LOCK_KEY
        self.conn_middle.set(self.train_lock_key, 0)
        self.episode = 0
        self.data_len = 0
        self.exit_flag = exit_flag
        self.log_loss_freq = config.LOG_LOSS_FREQUENCY
        self.timeout = config.TIME_OUT
        self.save_model = config.SAVE_MODELS
        self.save_freq = config.SAVE_FREQUENCY
        self.save_model_path = config.SAVE_MODEL_PATH
        self.model2redis_freq = config_nn.MODEL_TO_REDIS_FREQUENCY
        self.mimic_start = config.MIMIC_START
        if self.mimic_start:
            raise ValueError("imitation pre-training (MIMIC_START) is outside the B200 hot path: use the reference thread")
        self.load_checkpoint_path = config.LOAD_CHECKPOINT_PATH if config.LOAD_CHECKPOINT else None
        self.load_checkpoint_start = config.LOAD_EPISODE if config.LOAD_CHECKPOINT else 0
        self.sync = config.SYNC
        self.test = config.TEST
        self._loss_dict = defaultdict(list)


class BackwardGetDataThread(BackwardThread):
    def get_train_data(self) -> bool:
        """backward.py:146-152 with the decode deferred to the device: pop, raise the train lock, queue the raw payload."""
        item = self.conn_train.brpop(self.data_key, timeout=self.timeout * 3)
        if item is None:
            return False
        self.conn_middle.set(self.train_lock_key, 1)
        self.training_data_queue.put(item[1])
        return True

    def run(self) -> None:
        while self.exit_flag.value == b"0":
            if not self.get_train_data():
                break


class BackwardTrainThread(BackwardThread):
    def update_envstats_logger(self, dict_logger):
        for k, v in dict_logger.items():
            self.logger_f.add(({"mean": v}, self.data_len), k)

    def train_once(self, *queue_args) -> int:
        """One pass of the reference's while-body (backward.py:183-214); returns the iterations run."""
        t0 = time.time()
        train_data, dict_logger = self.training_data_queue.get(self.min_batch_size, *queue_args)
        train_data.to_tensor(dtype=self.tensortype, device=self.device)       # no-op for the device-decoded batch
        self.data_len += len(train_data)
        print("get training data ", time.time() - t0, flush=True)
        self.update_envstats_logger(dict_logger)
        iters = 0
        if not self.test:
            for loss_items, update_time, last in self.net.learn(train_data):
                update_time += self.load_checkpoint_start
                iters += 1
                for k, v in loss_items.items():
                    self._loss_dict[k].append(v)
                if last and update_time % self.model2redis_freq == 0:
                    self.net.nn2redis(self.pipe_middle, self.update_tag)
                if update_time % self.log_loss_freq == 0:
                    for key, vals in self._loss_dict.items():
                        if len(vals):
                            self.logger_f.add((sum(vals) / len(vals), update_time), key)
                            vals.clear()
                if last and self.save_model and update_time % self.save_freq == 0:
                    torch.save(self.net.state_dict(), self.save_model_path + "_" + str(update_time) + ".pt")
            self.pipe_middle.set(self.train_lock_key, 0)
            self.pipe_middle.execute()
        print("once backward COSTS: ", time.time() - t0, flush=True)
        return iters

    def prepare(self):
        if self.load_checkpoint_path:
            self.net.load_state_dict(torch.load(self.load_checkpoint_path))
        self.net.nn2redis(self.pipe_middle, self.update_tag)

    def run(self):
        self.prepare()
        while self.exit_flag.value == b"0":
            try:
                self.train_once(True, self.timeout * 3)
            except _queue.Empty:
                break
        print("backward exit !", flush=True)
