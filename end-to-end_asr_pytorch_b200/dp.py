"""Data-parallel training plumbing: one process per GPU, gradients only - over NVLink / NVSwitch peer memory with this
package's own all-reduce kernel (`PeerAllReduce`, csrc/allreduce.cu) where the box offers symmetric memory, over NCCL
otherwise.

The hot-path kernels have no cross-utterance coupling: every rank runs
them on its own utterances and the only exchange per step is the parameter-gradient
all-reduce.  `GradAllReduce` does that with bucket views (gradients accumulate
directly inside flat per-bucket buffers, no copies) and launches each bucket's
all-reduce from a post-accumulate hook as soon as its last gradient is written, in
reverse parameter order, so the collectives overlap the rest of the backward pass.
The reference itself is single-GPU (CIF.sh:5,71); this is new plumbing around it.

Loss normalisation note: `ctc_loss`, the quantity loss and the CE are local-batch
means; averaging gradients over ranks equals the global-batch gradient when every
rank holds the same number of utterances / tokens (true for the synthetic configs).
"""
import ctypes
import os

import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


class GradAllReduce:
    """Bucketed, overlapped gradient averaging for `module` (call `finish()` before the optimiser step).

        sync = GradAllReduce(model, bucket_mb=25)
        loss.backward()        # buckets are all-reduced as they fill
        sync.finish()          # wait + average; then optimizer.step()

    One backward pass per reset()/finish() pair (a second one raises: it would add local gradients to sums that
    are already reduced).  Collectives are issued in a fixed bucket order on every rank.
    """

    def __init__(self, module, bucket_mb=25.0, process_group=None, overlap=True, backend="auto"):
        """overlap=False registers no hooks: every bucket is all-reduced from finish().  That is the mode for a backward
        pass replayed from a CUDA graph (hooks only run while the graph is being captured, not when it is replayed).
        backend: "peer" = the buckets are ranges of ONE symmetric-memory buffer, averaged by this package's kernel over
        NVLink peer memory (PeerAllReduce; fp32 parameters on one device); "nccl" = one NCCL all-reduce per bucket;
        "auto" = peer when the process group, the device and the parameters allow it."""
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params
        self.buckets = []          # (flat buffer, [params])
        self._owner = {}
        cap = int(bucket_mb * 1024 * 1024)
        uniform = bool(params) and all(p.dtype == torch.float32 and p.device == params[0].device and p.is_cuda for p in params)
        auto = backend == "auto"
        if auto:
            backend = "peer" if (self.world > 1 and uniform and PeerAllReduce.available(params[0].device)) else "nccl"
        if backend == "peer" and not uniform:
            raise ValueError("GradAllReduce(backend='peer') needs fp32 parameters on one CUDA device")
        self.backend = backend
        groups = []
        cur, cur_bytes = [], 0
        # reverse order: the last layers' gradients are ready first
        for p in reversed(params):
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > cap or p.dtype != cur[0].dtype or p.device != cur[0].device):
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            groups.append(cur)
        self.peer = None
        self._ranges = []          # peer backend: (offset, padded length) of every bucket inside the symmetric buffer
        if backend == "peer":
            pad4 = lambda n: (n + 3) // 4 * 4
            total = sum(pad4(sum(p.numel() for p in g)) for g in groups)
            try:
                self.peer = PeerAllReduce(total, params[0].device, process_group)
            except Exception:           # noqa: BLE001  (no peer access on this box: every rank fails alike)
                if auto:
                    backend = self.backend = "nccl"
                else:
                    raise
        if backend == "peer":
            off = 0
            for g in groups:
                n = sum(p.numel() for p in g)
                self._ranges.append((off, pad4(n)))
                self._seal(g, self.peer.flat[off:off + n])
                off += pad4(n)
        else:
            for g in groups:
                self._seal(g)
        self._pending = [0] * len(self.buckets)
        self._handles = []
        self._next = 0             # buckets are all-reduced strictly in index order on every rank
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params] if overlap else []
        self.reset()

    def _seal(self, plist, flat=None):
        if flat is None:
            flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
        off = 0
        for p in plist:
            p.grad = flat[off:off + p.numel()].view_as(p)      # gradient lives inside the bucket
            self._owner[p] = len(self.buckets)
            off += p.numel()
        self.buckets.append((flat, plist))

    def reset(self):
        """Zero the buckets (instead of optimizer.zero_grad(), which would detach the views)."""
        for i, (flat, plist) in enumerate(self.buckets):
            flat.zero_()
            self._pending[i] = len(plist)
        self._handles = []
        self._next = 0

    def _launch_ready(self, force=False):
        """All-reduce buckets in index order: bucket i goes out once it is complete AND every bucket before it has gone
        out, so all ranks issue the same sequence of collectives even when their gradients become ready in a
        different order (or some never do: `force` sends those from finish())."""
        while self._next < len(self.buckets) and (force or self._pending[self._next] == 0):
            flat = self.buckets[self._next][0]
            if self.world > 1:
                if self.peer is not None:
                    if force and self._next == 0:
                        # nothing has gone out yet (graph replay, or finish() right after backward): one launch for all
                        self.peer.launch()
                        self._next = len(self.buckets)
                        self._handles.append((None, None))
                        return
                    self.peer.launch(*self._ranges[self._next])
                    self._handles.append((None, None))
                else:
                    self._handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat))
            self._next += 1

    def _on_grad(self, p):
        i = self._owner[p]
        self._pending[i] -= 1
        if self._pending[i] < 0 or i < self._next:
            # a second backward() into a bucket that may already have been summed over ranks would be reduced twice
            raise RuntimeError("GradAllReduce: a parameter received a second gradient before finish(); one backward pass "
                               "per reset()/finish() (accumulate micro-batches into the loss, or call finish() in between)")
        if self._pending[i] == 0:
            self._launch_ready()

    def finish(self):
        """Wait for the outstanding all-reduces and turn the sums into means."""
        self._launch_ready(force=True)      # buckets with a parameter that got no gradient this step go out here
        if self.peer is not None:
            if self._handles:
                self.peer.wait()           # the kernel writes means: nothing to divide
        else:
            for h, flat in self._handles:
                h.wait()
                flat.div_(self.world)
        for i, (_, plist) in enumerate(self.buckets):
            self._pending[i] = len(plist)
        self._handles = []
        self._next = 0

    def grad_bytes(self):
        return sum(flat.numel() * flat.element_size() for flat, _ in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()


class PeerAllReduce:
    """One flat fp32 gradient buffer in SYMMETRIC memory, averaged over the ranks by this package's kernel
    (`asr_allreduce_mean_f32`: NVSwitch multicast load-reduce + broadcast store, or plain peer loads / stores when the box
    has no multicast objects).  torch.distributed._symmetric_memory is used for what it is: the allocator and the exchange
    of the peer / multicast handles; no collective of torch or NCCL runs on the data.

        ar = PeerAllReduce(n_floats, device)          # collective: every rank of the group calls it
        ar.flat[...]                                  # the bucket (gradients accumulate here)
        ar.launch()                                   # on a side stream behind the current one; returns at once
        ar.wait()                                     # the current stream waits for the reduced values

    `available()` tells whether this process group / device can do it; GradAllReduce falls back to NCCL otherwise.
    """

    CTAS = 32

    @staticmethod
    def available(device=None):
        if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
            return False
        if dist.get_backend() != "nccl":
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
        except Exception:
            return False
        return True

    def __init__(self, numel, device, process_group=None, ctas=None, multicast=None):
        """multicast: False / None = peer loads and stores, True = the NVSwitch multicast flavour.  Bytes per link and
        direction for n gradient bytes: multicast n (1 + 1/W), peer 2 n (W - 1) / W.  Measured on B200 (209 MB): at two
        ranks the peer flavour wins outright (0.39 vs 0.60 ms); at eight the multicast flavour is faster alone (0.52 vs
        0.65 ms; NCCL 0.60) but disturbs the HBM-bound kernels it overlaps with more (hot-path step 3.23 vs 3.10 ms), so the
        peer flavour is the default.  ASR_ALLREDUCE_MULTICAST=0/1 overrides None."""
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.group = process_group if process_group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.ctas = int(ctas or self.CTAS)
        self.numel = int(numel)
        padded = (self.numel + 3) // 4 * 4
        flag_words = _lib.lib().asr_allreduce_signal_bytes(self.world, self.ctas) // 4
        with torch.cuda.device(device):
            self._buf = symm_mem.empty(padded, dtype=torch.float32, device=device)
            self._flags = symm_mem.empty(max(flag_words, 4), dtype=torch.int32, device=device)
            self._buf.zero_()
            self._flags.zero_()
            torch.cuda.synchronize()
            self._h = symm_mem.rendezvous(self._buf, self.group)
            self._hf = symm_mem.rendezvous(self._flags, self.group)
            self.stream = torch.cuda.Stream(device=device)
        self.flat = self._buf[:self.numel]
        self.multicast = int(self._h.multicast_ptr or 0)
        if multicast is None and os.environ.get("ASR_ALLREDUCE_MULTICAST") in ("0", "1"):     # measurement override
            multicast = os.environ["ASR_ALLREDUCE_MULTICAST"] == "1"
        if not multicast:
            self.multicast = 0
        if multicast is True and not self.multicast:
            raise RuntimeError("PeerAllReduce: this box offers no multicast address for symmetric memory")
        dist.barrier(self.group)            # every rank's flags are zero before anybody's first launch raises one
        self._done = torch.cuda.Event()

    def launch(self, offset=0, numel=None):
        """Average elements [offset, offset + numel) (multiples of 4) over the ranks, on the side stream, behind everything the
        current stream has queued so far."""
        n = (self.numel + 3) // 4 * 4 - offset if numel is None else int(numel)
        cur = torch.cuda.current_stream(self._buf.device)
        self.stream.wait_stream(cur)
        with torch.cuda.device(self._buf.device):
            self._lib.check(self._lib.lib().asr_allreduce_mean_f32(
                ctypes.c_void_p(int(self._h.buffer_ptrs_dev)), ctypes.c_void_p(self.multicast or None),
                ctypes.c_void_p(int(self._hf.buffer_ptrs_dev)), self.rank, self.world, int(offset), n, self.ctas,
                ctypes.c_void_p(self.stream.cuda_stream)), "asr_allreduce_mean_f32")
        self._done.record(self.stream)

    def wait(self):
        torch.cuda.current_stream(self._buf.device).wait_event(self._done)

    def flavour(self):
        return "nvswitch multicast (multimem.ld_reduce / multimem.st)" if self.multicast else "peer loads / stores"


def shard_utterances(n_utts, rank=None, world=None):
    """Contiguous utterance range of this rank (data parallel by utterance)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    per = (n_utts + world - 1) // world
    lo = min(n_utts, rank * per)
    return lo, min(n_utts, lo + per)
