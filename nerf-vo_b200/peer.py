"""Peer-visible device buffers for the fused gradient exchange (csrc/exchange.cu): plain cudaMalloc allocations exported
with CUDA IPC, opened by every other rank of the node, wrapped as torch tensors (PyTorch is the plumbing: the handles travel
through `torch.distributed.all_gather_object`).

    bufs = PeerBuffers(n_floats)            # collective: every rank of the default group calls it
    bufs.params / bufs.grads                # this rank's flat fp32 tensors (peer-mapped)
    bufs.adam_exchange_step(...)            # reduce-scatter + Adam + all-gather in one kernel

No fallback: if peer access is unavailable the constructor raises (use MappingTrainer(exchange="nccl") instead)."""
from __future__ import annotations

import ctypes
from typing import List

import torch
import torch.distributed as dist

from . import _lib


class _RawCudaBuffer:
    """Owner of one nvo_peer_alloc allocation, exposing __cuda_array_interface__ so torch can alias it."""

    def __init__(self, n_items: int, typestr: str, itemsize: int):
        ptr = ctypes.c_void_p()
        self.handle = (ctypes.c_ubyte * 64)()
        rc = _lib.load().nvo_peer_alloc(n_items * itemsize, ctypes.addressof(ptr), ctypes.addressof(self.handle))
        if rc != 0:
            raise RuntimeError(f"nvo_peer_alloc failed: {_lib.load().nvo_last_error().decode()}")
        self.ptr = int(ptr.value)
        self.__cuda_array_interface__ = {"shape": (n_items,), "typestr": typestr, "data": (self.ptr, False), "version": 2}

    def __del__(self):
        try:
            if getattr(self, "ptr", 0):
                _lib.load().nvo_peer_free(self.ptr)
                self.ptr = 0
        except Exception:
            pass


def slice_range(n: int, rank: int, world: int):
    """[lo, hi) in floats of the parameter slice rank `rank` owns (float4-aligned; mirrors nvo_exchange_slice)."""
    n4 = (n + 3) // 4
    chunk = (n4 + world - 1) // world
    a, b = min(rank * chunk, n4), min((rank + 1) * chunk, n4)
    return a * 4, b * 4


class PeerBuffers:
    def __init__(self, n_floats: int, device: torch.device, group=None):
        if n_floats % 4 != 0:
            raise RuntimeError("flat buffer size must be a multiple of 4 floats")
        lib = _lib.load()
        have_group = dist.is_available() and dist.is_initialized()
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if have_group else (0, 1)
        self.n = int(n_floats)
        self.device = device
        words = int(lib.nvo_exchange_flag_words())
        with torch.cuda.device(device):
            self._raw = [_RawCudaBuffer(self.n, "<f4", 4), _RawCudaBuffer(self.n, "<f4", 4), _RawCudaBuffer(words, "<i4", 4)]
            self.params = torch.as_tensor(self._raw[0], device=device)
            self.grads = torch.as_tensor(self._raw[1], device=device)
            self.flags = torch.as_tensor(self._raw[2], device=device)
            assert self.params.data_ptr() == self._raw[0].ptr and self.grads.data_ptr() == self._raw[1].ptr
            torch.cuda.synchronize()
            mine = [bytes(r.handle) for r in self._raw]
            everyone: List = [None] * self.world
            if have_group:
                dist.all_gather_object(everyone, mine, group=group)
            else:
                everyone[0] = mine
            self._opened: List[int] = []
            ptrs = [[0] * self.world for _ in range(3)]
            for k in range(self.world):
                for j in range(3):
                    if k == self.rank:
                        ptrs[j][k] = self._raw[j].ptr
                    else:
                        out = ctypes.c_void_p()
                        h = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[k][j])
                        rc = lib.nvo_peer_open(ctypes.addressof(h), ctypes.addressof(out))
                        if rc != 0:
                            raise RuntimeError(f"nvo_peer_open(rank {k}) failed: {lib.nvo_last_error().decode()}")
                        ptrs[j][k] = int(out.value)
                        self._opened.append(int(out.value))
            self._h_params = (ctypes.c_void_p * self.world)(*ptrs[0])
            self._h_grads = (ctypes.c_void_p * self.world)(*ptrs[1])
            self._h_flags = (ctypes.c_void_p * self.world)(*ptrs[2])
            lo, hi = slice_range(self.n, self.rank, self.world)
            self.slice = (lo, hi)
            chunk = slice_range(self.n, 0, self.world)[1]
            self.exp_avg = torch.zeros(max(chunk, 4), dtype=torch.float32, device=device)
            self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        if have_group:
            dist.barrier(group=group)  # every rank has mapped every buffer before anyone launches

    def adam_exchange_step(self, step: torch.Tensor, lr: float, beta1: float, beta2: float, eps: float) -> None:
        """One launch on the current stream: gradients of all ranks summed -> mean -> Adam on the own slice -> new parameters
        written into every rank's replica.  `step` is the device int32 counter (identical on every rank), incremented here."""
        _lib.call("nvo_adam_exchange_step", self.n, self.rank, self.world, ctypes.addressof(self._h_params), ctypes.addressof(self._h_grads),
                  ctypes.addressof(self._h_flags), self.exp_avg, self.exp_avg_sq, step, lr, beta1, beta2, eps, 1.0 / self.world)

    def add_group(self, offset: int, n: int) -> int:
        """Registers the flat range [offset, offset + n) as one optimizer parameter group: its own Adam moments (only this rank's slice of
        the range) and its own block of flags (exchange phase).  Returns the group id for adam_exchange_group."""
        if offset % 4 or n % 4 or n <= 0 or offset < 0 or offset + n > self.n:
            raise RuntimeError(f"parameter group [{offset}, +{n}) must be float4-aligned and inside the flat buffer of {self.n} floats")
        if not hasattr(self, "_groups"):
            self._groups = []
        if len(self._groups) >= 3:
            raise RuntimeError("at most 3 exchange groups (flag phases)")
        chunk = slice_range(n, 0, self.world)[1]
        with torch.cuda.device(self.device):
            m = torch.zeros(max(chunk, 4), dtype=torch.float32, device=self.device)
            self._groups.append((int(offset), int(n), m, torch.zeros_like(m)))
        return len(self._groups) - 1

    def group_moments(self, gid: int):
        return self._groups[gid][2], self._groups[gid][3]

    def adam_exchange_group(self, gid: int, step: torch.Tensor, lr: float, beta1: float, beta2: float, eps: float, ctas_per_sm: int = 0) -> None:
        """adam_exchange_step restricted to one registered group; `step` is the GROUP's device counter (incremented here).
        ctas_per_sm > 0 caps the persistent grid (for a launch that runs next to other kernels)."""
        off, n, m, v = self._groups[gid]
        _lib.call("nvo_adam_exchange_group", off, n, gid, self.rank, self.world, ctypes.addressof(self._h_params), ctypes.addressof(self._h_grads),
                  ctypes.addressof(self._h_flags), m, v, step, lr, beta1, beta2, eps, 1.0 / self.world, int(ctas_per_sm))

    def adam_exchange_group_decay(self, gid: int, step: torch.Tensor, lr_init: float, lr_final: float, max_steps: int, beta1: float, beta2: float,
                                  eps: float, ctas_per_sm: int = 0) -> None:
        """adam_exchange_group with ExponentialDecayScheduler's learning rate evaluated on the device (the "camera_opt" group)."""
        off, n, m, v = self._groups[gid]
        _lib.call("nvo_adam_exchange_group_decay", off, n, gid, self.rank, self.world, ctypes.addressof(self._h_params), ctypes.addressof(self._h_grads),
                  ctypes.addressof(self._h_flags), m, v, step, lr_init, lr_final, int(max_steps), beta1, beta2, eps, 1.0 / self.world, int(ctas_per_sm))

    def adam_exchange_groups2(self, gid_a: int, step_a: torch.Tensor, gid_b: int, step_b: torch.Tensor, lr: float, beta1: float, beta2: float, eps: float) -> None:
        """Both groups in ONE launch (one pair of barriers): for steps on which their gradients are complete at the same time."""
        oa, na, ma, va = self._groups[gid_a]
        ob, nb, mb, vb = self._groups[gid_b]
        _lib.call("nvo_adam_exchange_groups2", oa, na, ma, va, step_a, ob, nb, mb, vb, step_b, self.rank, self.world, ctypes.addressof(self._h_params),
                  ctypes.addressof(self._h_grads), ctypes.addressof(self._h_flags), lr, beta1, beta2, eps, 1.0 / self.world)

    def reset_epochs(self) -> None:
        """COLLECTIVE.  Zeroes this rank's flag pads.  The exchange kernels' barriers wait for `flag >= epoch` with the epoch taken from the
        group's device step counter, so whenever a counter is set BACK (the state restore behind MappingTrainer.capture()'s warm-up steps, a
        checkpoint load) flags left at a later epoch would let the next barriers pass before the peers have arrived: reduce-scatter over
        gradients that are still being written, identical replicas, wrong parameters.  Call it with no exchange in flight."""
        torch.cuda.synchronize(self.device)
        dist.barrier()
        self.flags.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier()

    def error_word(self) -> int:
        """0 = healthy; 1 / 2 = a peer never signalled 'gradients ready' / 'replicas written' (bounded spin timed out) in any phase."""
        words = int(_lib.load().nvo_exchange_flag_words())
        f = self.flags.cpu()
        return max([int(f[40 * ph + 33]) for ph in range(3) if 40 * ph + 33 < words] or [0])

    def close(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.nvo_peer_close(p)
        self._opened = []
