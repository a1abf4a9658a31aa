"""Momentum-SGD over ONE flat parameter / gradient / momentum buffer (the reference's optim.SGD of
networks/kd_model.py:74-75: momentum 0.9, weight decay on every parameter, lr written to param_groups[0]['lr']).

Parameters and their .grad become views into the flat buffers (layouts preserved), so
  * G_solver.step() is one fused kernel (csrc/pool.cu: sgd_kernel) instead of a per-tensor foreach,
  * zero_grad() is one memset,
  * the data-parallel exchange is ONE NCCL all-reduce of the flat gradient buffer; the 1/world averaging is folded into
    the update kernel (grad_scale).
"""
import torch
import torch.distributed as dist

from . import ops


def _dense(t):
    """non-overlapping and dense: some permutation of the dims is contiguous"""
    dims = sorted([(st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1], reverse=True)
    expect = 1
    for st, sz in reversed(dims):
        if st != expect:
            return False
        expect *= sz
    return True


def owned_range(total, world, rank):
    """[lo, hi) of a flat buffer of `total` floats (a multiple of 4) that `rank` reduces, updates and multicasts in step_nvls():
    contiguous, float4-aligned, the ranges of all ranks tile [0, total) exactly."""
    n4 = total // 4
    per = (n4 + world - 1) // world
    return 4 * min(rank * per, n4), 4 * min((rank + 1) * per, n4)


class FlatSGD:
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, direct_grads=False, symmetric=False):
        """direct_grads: the backward kernels of functions.py write each parameter's gradient straight into its view of the flat
        buffer (`p._skd_grad`) instead of returning a tensor for autograd to add -- one elementwise launch and one temporary less
        per parameter.  Valid because zero_grad() precedes every backward and no parameter is used twice in one forward; a caller
        that accumulates several backward passes into .grad must leave it off.
        symmetric: allocate the flat parameter and gradient buffers in torch symmetric memory (same virtual layout on every rank,
        NVSwitch multicast mapping) -- the precondition of `enable_nvls()`.  Collective: every rank must construct its optimizers
        in the same order."""
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        self.param_groups = [dict(params=self.params, lr=lr, initial_lr=lr, momentum=momentum, weight_decay=weight_decay)]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4                       # keep every tensor 16-byte aligned
        self._nvls = None
        if symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.flat_p = symm_mem.empty(total, dtype=torch.float32, device=dev); self.flat_p.zero_()
            self.flat_g = symm_mem.empty(total, dtype=torch.float32, device=dev); self.flat_g.zero_()
        else:
            self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
            self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self._views = []
        for p, off in zip(self.params, offs):
            src = p.data
            if src.dtype != torch.float32 or not _dense(src):
                raise ValueError("FlatSGD needs dense fp32 parameters")
            dst = torch.as_strided(self.flat_p, src.shape, src.stride(), off)
            dst.copy_(src)
            p.data = dst
            g = torch.as_strided(self.flat_g, src.shape, src.stride(), off)
            p.grad = g
            self._views.append(g)
            if direct_grads:
                p._skd_grad = g
        self.lr_dev = torch.tensor(float(lr), device=dev, dtype=torch.float32)
        self.steps = 0
        self.grad_scale = 1.0
        self._offsets = offs + [total]
        self._buckets = None
        self.producer_streams = []

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p, g in zip(self.params, self._views):
            if p.grad is not g:                                      # somebody detached it: re-attach the flat view
                p.grad = g

    def all_reduce_grads(self, world):
        """utils/parallel.py:54-63,155 semantics: mean over ranks of per-rank gradients."""
        if self._nvls is not None:                                   # the reduction happens inside step_nvls()
            return
        if world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            self.grad_scale = 1.0 / world
        else:
            self.grad_scale = 1.0

    # ---- gradient exchange fused into the update: NVSwitch multicast (multimem.ld_reduce / multimem.st) -----------------------------
    def enable_nvls(self, group=None):
        """Rendezvous of the symmetric flat buffers (collective).  Afterwards `step_nvls()` replaces all-reduce + step: this rank
        updates elements [lo, hi) -- 1/world of the parameters -- from the switch-reduced gradient and multicasts the result.
        Raises when the buffers are not symmetric or the fabric has no multicast support (caller falls back to NCCL)."""
        import os
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        if dist.get_world_size(group) > 2 and os.environ.get("SKD_NVLS_ANY_WORLD") != "1":
            # validated on hardware for 2 ranks only (tests/test_ddp_nccl_gpu.py); the one 4-rank run of this round did not finish
            # (DESIGN.md section 5).  Fail here, loudly, rather than risk a cross-rank hang inside a training job.
            raise RuntimeError("allreduce='nvls' is validated for world_size 2 only; use allreduce='nccl' (or set SKD_NVLS_ANY_WORLD=1 to try)")
        hp = symm_mem.rendezvous(self.flat_p, group)
        hg = symm_mem.rendezvous(self.flat_g, group)
        if not hp.multicast_ptr or not hg.multicast_ptr:
            raise RuntimeError("symmetric memory without multicast (NVLS) support on this system")
        world, rank = hp.world_size, hp.rank
        lo, hi = owned_range(self._offsets[-1], world, rank)
        self._nvls = dict(hp=hp, hg=hg, lo=lo, hi=hi, world=world, rank=rank)
        self.grad_scale = 1.0 / world

    def step_nvls(self):
        """barrier | ld_reduce(grad) -> SGD on the owned range -> multicast store(param) | barrier, on the current stream (capturable).
        The first barrier orders every rank's gradient writes before any rank's reduction; the second keeps every rank from reading
        parameters -- or zeroing gradients -- before all ranges have been reduced and broadcast."""
        nv = self._nvls
        g = self.param_groups[0]
        self.lr_dev.fill_(float(g['lr']))
        nv["hg"].barrier(channel=0)
        ops.sgd_step_nvls(nv["lo"], nv["hi"], nv["hp"].multicast_ptr, nv["hg"].multicast_ptr, self.flat_p, self.flat_m, self.lr_dev,
                          g['momentum'], g['weight_decay'], 1.0 / nv["world"])
        nv["hp"].barrier(channel=1)
        self.steps += 1

    # ---- all-reduce overlapped with the backward pass -------------------------------------------------------------------------
    def enable_overlap(self, n_buckets=4):
        """Split the flat gradient into `n_buckets` contiguous ranges of whole parameters (forward order).  During backward a
        post-accumulate hook counts the parameters of each range; when a range is complete its NCCL all-reduce is issued on a side
        stream -- the ranges of the late layers (finished first) travel while the early layers are still being differentiated.
        CUDA-graph capturable: the side stream forks from / joins the capturing stream."""
        n = len(self.params)
        total = self._offsets[-1]
        bounds, target, nxt = [0], total / n_buckets, 1
        for i in range(n):
            if self._offsets[i + 1] >= target * nxt and len(bounds) < n_buckets:
                bounds.append(i + 1); nxt += 1
        if bounds[-1] != n:
            bounds.append(n)
        self._buckets = [dict(lo=lo, hi=hi, start=self._offsets[lo], end=self._offsets[hi], left=0, sent=False)
                         for lo, hi in zip(bounds[:-1], bounds[1:]) if hi > lo]
        which = {}
        for bi, b in enumerate(self._buckets):
            for i in range(b["lo"], b["hi"]):
                which[i] = bi
        self._side = torch.cuda.Stream()
        self._active = False
        self._arrival_log = None                                     # debugging aid: list -> (param index, via) per arrival
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, bi=which[i], i=i: self._arrived(bi, i, "hook"))
            p._skd_arrived = (lambda bi=which[i], i=i: self._arrived(bi, i, "direct"))   # called by backward kernels that wrote p._skd_grad directly

    def begin_overlapped_reduce(self, world):
        self._world = world
        self.grad_scale = 1.0 / world if world > 1 else 1.0
        for b in self._buckets:
            b["left"], b["sent"] = b["hi"] - b["lo"], False
        self._got = bytearray(len(self.params))
        self._active = world > 1

    def _send(self, b):
        b["sent"] = True
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)                                  # the bucket's gradients are complete on the compute stream
        for s in self.producer_streams:                              # ... and on the streams that write gradients besides it (wgrad overlap)
            self._side.wait_stream(s)
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.flat_g[b["start"]:b["end"]], op=dist.ReduceOp.SUM)

    def _arrived(self, bi, i=None, via=None):
        if self._arrival_log is not None:
            self._arrival_log.append((i, via))
        if not self._active:
            return
        # a parameter counts ONCE per pass: a gradient the kernels wrote directly reports itself (functions._grad_written) and the
        # autograd engine may still run the parameter's post-accumulate hook for the `None` the backward returned (torch 2.11 does)
        if i is not None:
            if self._got[i]:
                return
            self._got[i] = 1
        b = self._buckets[bi]
        b["left"] -= 1
        if b["left"] == 0 and not b["sent"]:
            self._send(b)

    def finish_overlapped_reduce(self):
        """Reduce whatever has not been sent (parameters that received no gradient) and join the side stream."""
        if not self._active:
            return
        for b in self._buckets:
            if not b["sent"]:
                self._send(b)
        torch.cuda.current_stream().wait_stream(self._side)
        self._active = False

    def step(self):
        if self._nvls is not None:
            return self.step_nvls()
        g = self.param_groups[0]
        self.lr_dev.fill_(float(g['lr']))
        # momentum buffer starts at zero, so the first step needs no special case (v = mu*0 + d)
        ops.sgd_step(self.flat_p, self.flat_g, self.flat_m, self.lr_dev, g['momentum'], g['weight_decay'], False, self.grad_scale)
        self.steps += 1

    def state_dict(self):
        m = self.flat_m.clone()
        if self._nvls is not None:                                   # each rank holds the momentum of its own range only (zeros elsewhere)
            m[:self._nvls["lo"]] = 0; m[self._nvls["hi"]:] = 0
            dist.all_reduce(m, op=dist.ReduceOp.SUM)
        return dict(momentum=m, steps=self.steps, lr=self.param_groups[0]['lr'])

    def load_state_dict(self, sd):
        self.flat_m.copy_(sd['momentum']); self.steps = sd['steps']; self.param_groups[0]['lr'] = sd['lr']
