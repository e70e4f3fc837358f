"""Rectified Adam -- drop-in for the ``RAdam`` class of the reference's ``radam.py`` (:5-94).

Same constructor, defaults, param-group keys (including the ``buffer`` cache, so optimizer state dicts
interchange with the reference's checkpoints) and per-parameter state (``step``, ``exp_avg``,
``exp_avg_sq``).  ``step()`` differs in execution only: parameters that sit back to back in one allocation
(the 16 hash-table levels; the five matrices of a NeRFSmall) and share hyper-parameters are updated by ONE
fused CUDA kernel launch (hn_radam_step) instead of ~10 ATen launches per tensor.
"""
from __future__ import annotations

import math

import torch
from torch.optim.optimizer import Optimizer

from hn_b200 import _lib, ops


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=False,
                 fused_zero_grad=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        self.degenerated_to_sgd = degenerated_to_sgd
        self.grad_scale = 1.0  # set to 1/world_size by the data-parallel wrapper after a summed all-reduce
        # opt-in (not in the reference's signature): step() also clears the gradients it consumed, in the same
        # pass over memory, so that the zero_grad() of the next iteration (run_nerf.py:612) has nothing left to
        # fill.  Observable difference: p.grad reads zero right after step() instead of the consumed gradient.
        self.fused_zero_grad = bool(fused_zero_grad)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                        buffer=[[None, None, None] for _ in range(10)])
        super().__init__(params, defaults)

    # The span plan caches parameter runs, group indices and moment pointers: anything that replaces param groups
    # or state (checkpoint resume, unpickling, a new group) must drop it, or later lr decay would be applied to a
    # stale group dict and the kernel would run over moments that are no longer one allocation.
    def _drop_plan(self):
        self.__dict__.pop("_span_cache", None)
        self.__dict__.pop("_g_spans", None)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._drop_plan()

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__.setdefault("grad_scale", 1.0)
        self.__dict__.setdefault("fused_zero_grad", False)
        self._drop_plan()

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._drop_plan()

    # radam.py:62-78 -- depends on the step count only
    def _rectification(self, step, beta1, beta2):
        beta2_t = beta2 ** step
        n_max = 2 / (1 - beta2) - 1
        n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
        if n_sma >= 5:
            size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) \
                / (1 - beta1 ** step)
            return 1, size
        if self.degenerated_to_sgd:
            return 2, 1.0 / (1 - beta1 ** step)
        return 0, -1.0

    def _init_state(self, run):
        """Zero moments for a run of memory-consecutive parameters, allocated as one flat buffer each."""
        n = sum(p.numel() for p in run)
        m = torch.zeros(n, dtype=torch.float32, device=run[0].device)
        v = torch.zeros(n, dtype=torch.float32, device=run[0].device)
        off = 0
        for p in run:
            st = self.state[p]
            st['step'] = 0
            st['exp_avg'] = m[off:off + p.numel()].view_as(p)
            st['exp_avg_sq'] = v[off:off + p.numel()].view_as(p)
            off += p.numel()

    def _spans(self):
        """[(group index, [params])]: maximal runs of parameters whose data, gradients and moments are each back
        to back in one storage and whose step counts agree -- each run is updated by one kernel launch.  The plan is
        cached and re-validated per step by the gradient pointers and the moment pointers of every parameter
        (gradients normally live in persistent buffers; all parameters of a cached span advance their step counts
        together, so the counts stay equal); it holds group INDICES, so a replaced group dict is picked up."""
        cached = getattr(self, "_span_cache", None)
        if cached is not None:
            sig, spans = cached
            if sig == self._signature():
                return spans
        spans = self._compute_spans()
        self._span_cache = (self._signature(), spans)
        return spans

    def _signature(self):
        sig = []
        for group in self.param_groups:
            for p in group['params']:
                g = p.grad
                st = self.state.get(p)
                m = st.get('exp_avg') if st else None
                v = st.get('exp_avg_sq') if st else None
                sig.append((id(p), None if g is None else g.data_ptr(), None if m is None else m.data_ptr(),
                            None if v is None else v.data_ptr()))
        return sig

    def _compute_spans(self):
        out = []
        for gi, group in enumerate(self.param_groups):
            active = [p for p in group['params'] if p.grad is not None]
            for p in active:
                if p.grad.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("hashnerf_b200 RAdam updates fp32 CUDA parameters only (no CPU fallback)")
            runs, cur = [], []
            for p in active:  # runs of parameters that are consecutive in memory
                if cur and ops._consecutive([cur[-1], p]):
                    cur.append(p)
                else:
                    if cur:
                        runs.append(cur)
                    cur = [p]
            if cur:
                runs.append(cur)
            for run in runs:
                fresh = [p for p in run if len(self.state[p]) == 0]
                if len(fresh) == len(run):
                    self._init_state(run)
                else:
                    for p in fresh:
                        self._init_state([p])
                for p in run:
                    st = self.state[p]
                    if st['exp_avg'].dtype != torch.float32 or not st['exp_avg'].is_contiguous():
                        st['exp_avg'] = st['exp_avg'].float().contiguous()
                        st['exp_avg_sq'] = st['exp_avg_sq'].float().contiguous()
                    if not p.grad.is_contiguous():
                        p.grad = p.grad.contiguous()
                i = 0
                while i < len(run):
                    j = i + 1
                    st0 = self.state[run[i]]
                    while j < len(run):
                        a, b = run[j - 1], run[j]
                        sa, sb = self.state[a], self.state[b]
                        if not (sb['step'] == st0['step'] and ops._consecutive([a.grad, b.grad])
                                and ops._consecutive([sa['exp_avg'], sb['exp_avg']])
                                and ops._consecutive([sa['exp_avg_sq'], sb['exp_avg_sq']])):
                            break
                        j += 1
                    out.append((gi, run[i:j]))
                    i = j
        return out

    def _advance(self, group, span):
        """Host side of one update: bump the step counters, return (mode, step_size) (radam.py:61-78)."""
        for p in span:
            self.state[p]['step'] += 1
        beta1, beta2 = group['betas']
        return self._rectification(self.state[span[0]]['step'], beta1, beta2)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        zero = 1 if self.fused_zero_grad else 0
        for gi, span in self._spans():
            group = self.param_groups[gi]
            mode, step_size = self._advance(group, span)
            beta1, beta2 = group['betas']
            first, st = span[0], self.state[span[0]]
            n = sum(p.numel() for p in span)
            with ops._on(first.device):
                _lib.call("hn_radam_step", first.data_ptr(), first.grad.data_ptr(), st['exp_avg'].data_ptr(),
                          st['exp_avg_sq'].data_ptr(), n, beta1, beta2, group['eps'], float(group['lr']),
                          float(group['weight_decay']), float(step_size), mode, self.grad_scale, zero, ops._stream())
            if zero:
                ops.GradSink.note_cleared(first.grad.data_ptr(), n)
        ops.param_epoch[0] += 1  # the kernel wrote the parameters through raw pointers (no torch version bump)
        return loss

    # ---- CUDA-graph support ------------------------------------------------------------------------
    # A captured graph replays kernel launches, not Python.  The step-dependent scalars therefore live in a
    # device array that the captured kernels read and that is refreshed, per step, by a plain stream-ordered copy
    # issued right before the replay:
    #     opt.graph_plan()            once, gradients present, before capture
    #     opt.graph_launch()          inside the capture (kernels only)
    #     opt.graph_prepare()         before every replay, on the replay's stream: counters, lr schedule ->
    #                                 one slot of a pinned ring -> async copy to the device array
    # The ring matters: the host runs ahead of the GPU, and a single pinned row overwritten for step k+1 before the
    # copy of step k has executed would hand step k the scalars of step k+1 (harmless-looking on one GPU, but it makes
    # data-parallel ranks, whose hosts run ahead by different amounts, drift apart bit by bit).  A slot is rewritten
    # only after the event recorded behind its previous copy has completed.
    _RING = 4

    @torch.no_grad()
    def graph_plan(self):
        self._g_spans = self._spans()
        dev = self._g_spans[0][1][0].device
        n = len(self._g_spans)
        self._g_host = [torch.zeros(n, 8, dtype=torch.float32).pin_memory() for _ in range(self._RING)]
        self._g_events = [None] * self._RING
        self._g_slot = 0
        self._g_dev = torch.zeros(n, 8, dtype=torch.float32, device=dev)

    def graph_prepare(self):
        slot = self._g_slot
        self._g_slot = (slot + 1) % self._RING
        if self._g_events[slot] is not None:
            self._g_events[slot].synchronize()       # the copy that last read this pinned slot is done
        host = self._g_host[slot]
        for i, (gi, span) in enumerate(self._g_spans):
            group = self.param_groups[gi]
            mode, step_size = self._advance(group, span)
            beta1, beta2 = group['betas']
            row = host[i]
            row[0], row[1], row[2] = beta1, beta2, group['eps']
            row[3] = group['weight_decay'] * group['lr']
            row[4] = step_size * group['lr']
            row[5], row[6] = self.grad_scale, float(mode)
            row[7] = 1.0 if self.fused_zero_grad else 0.0
        with torch.cuda.device(self._g_dev.device):
            self._g_dev.copy_(host, non_blocking=True)   # stream-ordered: lands before the replay's kernels
            ev = self._g_events[slot] or torch.cuda.Event()
            ev.record()
            self._g_events[slot] = ev

    @torch.no_grad()
    def graph_launch(self):
        for i, (_gi, span) in enumerate(self._g_spans):
            first, st = span[0], self.state[span[0]]
            n = sum(p.numel() for p in span)
            with ops._on(first.device):
                _lib.call("hn_radam_step_dev", first.data_ptr(), first.grad.data_ptr(), st['exp_avg'].data_ptr(),
                          st['exp_avg_sq'].data_ptr(), n, self._g_dev[i].data_ptr(), ops._stream())
