"""DiffusionJumpySampler / ModelAdapter -- B200 host mirror of sampler/jumpy_sampler.py:54-307.

Same constructor kwargs, attributes, `sample(cond_c, seq_len, init)` and `get_sampling_info()`.
Each jump is: decoder forward (the caller's module, out of scope) -> ONE fused kernel
(`fddm_jump_step`: softmax -> Delta-step posterior or alpha-bar mix -> argmax / exponential-race
resampling, optionally emitting p_x0 and its argmax) preceded in exact mode by the tiny sync-free
coefficient kernel.  No host synchronisation happens inside `sample`.

Reference quirks that are reproduced on purpose: Q2 exact mode feeds the
T_infer-axis t to the T_train-length beta table; Q3 fast mode's 0-based/1-based alpha-bar index;
Q7 any posterior_mode other than "max" goes through `_to_indices`; Q8 `init` is ignored;
Q9 the ids returned are argmax of the LAST p_x0, the last jump's resampled ids are dropped.
"""
from __future__ import annotations

from typing import Literal, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from .scheduler import multistep_coeffs, philox_seed_offset


class ModelAdapter:
    """Unifies the denoising decoder as predict_x0_logits(x_t_idx, t, cond_c).  sampler:54-83."""

    def __init__(self, decoder):
        self.decoder = decoder

    @torch.no_grad()
    def predict_x0_logits(self, x_t_idx: Tensor, t: Tensor, cond_c: Tensor) -> Tensor:
        return self.decoder(x_t_idx, t, cond_c)               # positional call, sampler:82


class DiffusionJumpySampler:
    def __init__(self, scheduler, decoder, K: int, T_train: int, T_infer: int, r: int = 2, greedy: bool = True,
                 posterior_mode: Literal["average", "max"] = "average",
                 sampling_mode: Literal["exact", "fast"] = "exact", temperature: float = 1.0,
                 device: Optional[torch.device] = None):
        self.scheduler = scheduler
        self.model = ModelAdapter(decoder)
        self.K = int(K)
        self.T_train = int(T_train)
        self.T_infer = int(T_infer)
        self.r = int(r)
        self.greedy = bool(greedy)
        self.posterior_mode = posterior_mode
        self.sampling_mode = sampling_mode
        self.temperature = float(temperature)
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        alpha_bar = getattr(self.scheduler, "alpha_bar", None)
        if alpha_bar is None:                                   # sampler:131-133
            raise ValueError("scheduler must provide alpha_bar")
        self.alpha_bar = torch.as_tensor(alpha_bar, dtype=torch.float32, device=self.device)
        # parity-test hook: callable(step_index, (B, L, K)) -> fp32 Exp(1) noise tensor, or None
        self.noise_fn = None
        self.generator: Optional[torch.Generator] = None
        # optional int64[2] device tensor {seed, offset} read by the kernel instead of the host-side
        # generator state; every sampling jump advances the offset on the device (a captured add), which
        # makes a jump -- or a whole chain -- replayable in a CUDA graph with fresh noise per replay
        self.philox_state: Optional[torch.Tensor] = None
        self._graph_enabled = False
        self._graphs = {}

    def enable_cuda_graph(self, enabled: bool = True) -> "DiffusionJumpySampler":
        """capture the whole jump chain of `sample()` -- every decoder forward, coefficient
        kernel and fused jump kernel -- in ONE CUDA graph per (cond shape, seq_len) and replay it on later calls.
        Nothing in the chain synchronises with the host, so the replay removes the per-jump launch and Python
        overhead that dominates at small batch (the reference's evaluation builds a sampler per utterance, B=1:
        models/evaluate.py:165-174).  Requirements: the decoder must be CUDA-graph capturable; the returned
        tensors are the graph's static outputs and are overwritten by the next `sample()` with the same shapes
        (clone them to keep them).  Sampling draws use the device-side Philox state, advanced inside the graph,
        so every replay draws fresh noise."""
        self._graph_enabled = bool(enabled)
        if not enabled:
            self._graphs = {}
        return self

    # sampler:219-236, including the 0-based table indexed by a 1-based train-axis index (Q3)
    def _alpha_bar_index(self, t_infer_scalar: int) -> int:
        if t_infer_scalar <= 0:
            return -1                                           # alpha-bar_0 := 1
        ratio = float(t_infer_scalar) / float(max(1, self.T_infer))
        t_train_float = max(1.0, min(float(self.T_train), ratio * float(self.T_train)))
        idx = int(round(t_train_float))
        if idx >= self.alpha_bar.numel():
            raise IndexError(f"index {idx} is out of bounds for dimension 0 with size {self.alpha_bar.numel()}")
        return idx

    def _alpha_bar_at_t_train(self, t_infer_scalar: int) -> Tensor:
        idx = self._alpha_bar_index(t_infer_scalar)
        if idx < 0:
            return torch.tensor(1.0, device=self.device, dtype=torch.float32)
        return self.alpha_bar[idx]

    @torch.no_grad()
    def _jump_once(self, x_t_idx: Tensor, t_scalar: int, delta: int, cond_c: Tensor, seq_len: int, *,
                   want_p: bool = True, step: int = 0) -> Tuple[Tensor, Tensor]:
        """sampler:167-217.  Returns (x_{t-delta} ids, p_x0).  With want_p=False p_x0 is not
        materialised (None is returned in its place) -- `sample` only needs the last one."""
        x_new, p_x0, _ = self._jump(x_t_idx, t_scalar, delta, cond_c, seq_len, want_p=want_p, want_argmax=False,
                                    step=step)
        return x_new, p_x0

    def _jump(self, x_t_idx, t_scalar, delta, cond_c, seq_len, *, want_p, want_argmax, step):
        device = x_t_idx.device
        B = x_t_idx.size(0)
        t_tensor = torch.full((B,), t_scalar, device=device, dtype=torch.long)
        logits = self.model.predict_x0_logits(x_t_idx, t_tensor, cond_c)      # [B, L, K]
        if logits.dim() != 3 or logits.size(0) != B or logits.size(-1) != self.K:
            raise ValueError(f"decoder returned logits of shape {tuple(logits.shape)}, expected ({B}, L, {self.K})")
        Lq = logits.size(1)
        if Lq != seq_len or tuple(x_t_idx.shape) != (B, Lq):
            raise ValueError("x_t_idx / logits / seq_len disagree on the sequence length")
        dev = L.require_cuda(logits, x_t_idx)
        dt = L.dtype_code(logits)
        logits = logits.contiguous()
        x_t = x_t_idx.long().contiguous()

        flags = 0
        coeffs = None
        abar_index = -1
        alpha_bar = None
        if self.sampling_mode == "exact":
            flags |= L.JUMP_EXACT
            # Q2: t on the T_infer axis indexes the scheduler's (T_train-length) beta table as is
            # (works with any scheduler object that has a `betas` table, e.g. the reference's own class)
            betas = getattr(self.scheduler, "betas", None)
            if betas is None:
                raise ValueError("scheduler must provide betas for sampling_mode='exact'")
            coeffs = multistep_coeffs(t_tensor, int(delta), betas, int(getattr(self.scheduler, "K", self.K)))
            eps = float(getattr(self.scheduler, "eps", 1e-8))
        else:
            abar_index = self._alpha_bar_index(max(0, t_scalar - delta))
            alpha_bar = self.alpha_bar if self.alpha_bar.device == dev else self.alpha_bar.to(dev)
            eps = 1e-8
        sample = (self.posterior_mode != "max") and (not self.greedy)          # sampler:212-215, 153-162
        noise = None
        seed = offset = 0
        if sample:
            flags |= L.JUMP_SAMPLE
            if self.noise_fn is not None:
                noise = self.noise_fn(step, (B, Lq, self.K))
            if noise is not None:
                noise = noise.to(device=dev, dtype=torch.float32).contiguous()
                if noise.numel() != B * Lq * self.K:
                    raise ValueError("injected noise must have B*L*K elements")
            elif self.philox_state is None:
                seed, offset = philox_seed_offset(dev, self.generator, 4)
        p_x0 = None
        if getattr(self, "_debug_weights", False) and sample and noise is None and self.temperature == 1.0:
            # test hook (FDDM_JUMP_DEBUG_W): p_x0's slot receives the normalised distribution the in-kernel-RNG
            # flavour draws from, so its fast arithmetic can be checked against the oracle entry by entry
            flags |= L.JUMP_DEBUG_W
            want_p = False
            p_x0 = torch.empty_like(logits)
        if want_p:
            flags |= L.JUMP_WRITE_P
            p_x0 = torch.empty_like(logits)                     # softmax output keeps the logits dtype
        amax = torch.empty_like(x_t) if want_argmax else None
        x_out = torch.empty_like(x_t)
        ws = L.zeroed_workspace(dev, "jump", int(L.lib.fddm_jump_workspace_bytes(B, Lq)))
        L.check(L.lib.fddm_jump_step(logits.data_ptr(), dt, x_t.data_ptr(), L.ptr(coeffs), L.ptr(alpha_bar),
                                     abar_index, B, Lq, self.K, flags, self.temperature, eps, L.ptr(noise), seed,
                                     offset, L.ptr(self.philox_state if (sample and noise is None) else None),
                                     ws.data_ptr(), x_out.data_ptr(), L.ptr(amax), L.ptr(p_x0),
                                     L.stream_ptr(dev)), "jump_step")
        if sample and noise is None and self.philox_state is not None:
            # stream-ordered (and graph-capturable) advance of the device-side offset: the next jump of the
            # chain -- and the next replay of a captured chain -- draws fresh variates
            self.philox_state[1:].add_(4)
        return x_out, p_x0, amax

    @torch.no_grad()
    def sample(self, cond_c: Tensor, seq_len: int, init: Literal["uniform", "random"] = "uniform", *,
               x_init: Optional[Tensor] = None, return_p: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
        """sampler:241-293.  Returns (x_0 ids [B,L], p_x0_last [B,L,K]).
        `x_init` (extra, keyword-only): the initial ids x_T instead of drawing them here (parity tests,
        CUDA-graph replay with caller-owned state).  `return_p=False` (extra): the last jump emits only
        argmax p_x0 -- the ids -- and p_x0_last is None (the evaluation callers drop it: evaluate.py:176,422)."""
        if self._graph_enabled and cond_c.is_cuda and not torch.cuda.is_current_stream_capturing():
            return self._sample_graphed(cond_c, seq_len, x_init, return_p)
        return self._sample_chain(cond_c, seq_len, x_init, return_p)

    @torch.no_grad()
    def sample_utterances(self, cond_c: Tensor, seq_len: int, num_samples: int = 1, *,
                          x_init: Optional[Tensor] = None, return_p: bool = False,
                          per_utterance_draws: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
        """the batched form of the reference's evaluation loops, which build one B=1
        sampler per utterance (models/evaluate.py:163-176) and, for multi-sample decoding, per utterance AND per
        sample (evaluate.py:405-422) -- B (x num_samples) chains of B=1 kernels.  Here all of them are rows of ONE
        chain: `cond_c[b]` is repeated `num_samples` times and every row has its own x_T and its own draws.

        Returns (ids [B, num_samples, L], p_x0 [B, num_samples, L, K] or None).  The jump kernels are
        row-independent, so with a row-independent decoder, greedy / "max" decoding gives, for every utterance,
        exactly the ids of its own B=1 chain started from the same x_T (GPU test
        `test_sample_utterances_equals_per_utterance_loop`).  `per_utterance_draws=True` draws x_T with one
        `randint((1, L))` per row, in the loop's order, so the generator is consumed draw for draw like the loop."""
        if cond_c.dim() < 1 or num_samples < 1:
            raise ValueError("sample_utterances: cond_c must have a batch axis and num_samples must be >= 1")
        B = cond_c.size(0)
        S = int(num_samples)
        rows = B * S
        cond_rows = cond_c if S == 1 else cond_c.repeat_interleave(S, dim=0)
        if x_init is not None:
            x_T = x_init.reshape(rows, seq_len).long()
        elif per_utterance_draws:
            x_T = torch.cat([torch.randint(low=0, high=self.K, size=(1, seq_len), device=cond_c.device,
                                           generator=self.generator) for _ in range(rows)], dim=0)
        else:
            x_T = torch.randint(low=0, high=self.K, size=(rows, seq_len), device=cond_c.device,
                                generator=self.generator)
        ids, p = self.sample(cond_rows, seq_len, x_init=x_T, return_p=return_p)
        return ids.view(B, S, seq_len), (p.view(B, S, seq_len, self.K) if p is not None else None)

    @torch.no_grad()
    def _sample_graphed(self, cond_c: Tensor, seq_len: int, x_init: Optional[Tensor],
                        return_p: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
        B = cond_c.size(0)
        dev = cond_c.device
        key = (tuple(cond_c.shape), cond_c.dtype, int(seq_len), dev.index, bool(return_p))
        # x_T is drawn outside the graph (one tiny kernel, same generator semantics as the eager path)
        x_T = x_init.long() if x_init is not None else torch.randint(low=0, high=self.K, size=(B, seq_len), device=dev,
                                                                      generator=self.generator)
        ent = self._graphs.get(key)
        if ent is None:
            sampling = (self.posterior_mode != "max") and (not self.greedy)
            if sampling and self.noise_fn is None and self.philox_state is None:
                seed, _ = philox_seed_offset(dev, self.generator, 4)
                self.philox_state = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
            s_cond = cond_c.clone()
            s_x = x_T.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                         # warm-up outside capture (lazy init, autotune)
                for _ in range(2):
                    self._sample_chain(s_cond, seq_len, s_x, return_p)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._sample_chain(s_cond, seq_len, s_x, return_p)
            ent = self._graphs[key] = (graph, s_cond, s_x, out)
        graph, s_cond, s_x, out = ent
        s_cond.copy_(cond_c)
        s_x.copy_(x_T)
        graph.replay()
        return out

    @torch.no_grad()
    def _sample_chain(self, cond_c: Tensor, seq_len: int, x_init: Optional[Tensor],
                      return_p: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
        B = cond_c.size(0)
        device = cond_c.device
        if x_init is not None:
            if tuple(x_init.shape) != (B, seq_len):
                raise ValueError(f"x_init must have shape ({B}, {seq_len})")
            x_t_idx = x_init.long()
        else:
            # both `init` values draw uniform random ids in the reference (Q8, sampler:276-280)
            x_t_idx = torch.randint(low=0, high=self.K, size=(B, seq_len), device=device, generator=self.generator)
        t = self.T_infer
        p_x0_last = None
        x_0_idx = None
        step = 0
        while t > 0:
            delta = min(self.r, t)
            last = (t - delta) <= 0
            # only the last p_x0 is returned, so only the last jump writes it; its argmax (the
            # sampler's final x_0, sampler:292) is fused into the same kernel
            x_t_idx, p, amax = self._jump(x_t_idx, t, delta, cond_c, seq_len, want_p=last and return_p,
                                          want_argmax=last, step=step)
            if last:
                p_x0_last, x_0_idx = p, amax
            t -= delta
            step += 1
        if x_0_idx is None:                                      # T_infer <= 0: the reference fails here too
            raise AttributeError("'NoneType' object has no attribute 'argmax'")
        self.last_resampled_idx = x_t_idx                         # dropped by the reference (Q9); kept for inspection
        return x_0_idx, p_x0_last

    def get_sampling_info(self) -> dict:
        return {
            "sampling_mode": self.sampling_mode,
            "posterior_mode": self.posterior_mode,
            "T_infer": self.T_infer,
            "r": self.r,
            "greedy": self.greedy,
            "temperature": self.temperature,
            "K": self.K,
        }
