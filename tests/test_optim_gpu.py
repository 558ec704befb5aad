"""FusedAdamW (clip_grad_norm_ + transformers.AdamW + bf16 shadow refresh, csrc/optim.cu) against a
torch restatement of what the reference trains with (src/train.py:137,194-197):

    torch.nn.utils.clip_grad_norm_(params, 1.0)          # coef = min(1, max_norm / (||g|| + 1e-6))
    transformers.AdamW(lr, betas=(.9,.999), eps=1e-6, weight_decay=0, correct_bias=True).step()
        exp_avg    = b1 exp_avg    + (1 - b1) g
        exp_avg_sq = b2 exp_avg_sq + (1 - b2) g^2
        p -= lr sqrt(1 - b2^t) / (1 - b1^t) * exp_avg / (sqrt(exp_avg_sq) + eps)   # eps OUTSIDE the correction

These kernels run inside bench.py's timed region. fp32 elementwise arithmetic: tolerance 2e-6
relative on the update, the bf16 shadow must equal bf16(master) exactly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _small_model(cuda, n_layer=2):
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.model import MMTG
    g2 = {"n_layer": n_layer}
    model = MMTG(model_cfgs, data_config(), 13317, train_flag=True, token_table=synth.make_token_table(), gpt2_config=g2)
    model.set_dropout(0.0, 0.0, 0.0)
    model.load_state_dict(synth.make_state_dict(3, gpt2_cfg=g2))
    model.to(cuda)
    host = synth.batch_to_torch(synth.make_batch(2, seed=5))
    return model, {k: v.to(cuda) for k, v in host.items()}


class _RefAdamW:
    """The reference's clip + HF AdamW restated with torch ops in float64 accumulators-free fp32
    (same operation order as transformers/optimization.py AdamW.step of the pinned 4.12.3)."""

    def __init__(self, P, lr, b1=0.9, b2=0.999, eps=1e-6, max_norm=1.0):
        self.p = P.clone()
        self.m, self.v = torch.zeros_like(P), torch.zeros_like(P)
        self.lr, self.b1, self.b2, self.eps, self.max_norm, self.t = lr, b1, b2, eps, max_norm, 0

    def step(self, g):
        g = g.clone()
        if self.max_norm is not None:
            total = torch.linalg.vector_norm(g.double()).float()  # clip_grad_norm_: norm of per-tensor norms
            coef = torch.clamp(self.max_norm / (total + 1e-6), max=1.0)
            g = g * coef
        self.t += 1
        self.m.mul_(self.b1).add_(g, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
        denom = self.v.sqrt().add_(self.eps)
        step_size = self.lr * (1 - self.b2 ** self.t) ** 0.5 / (1 - self.b1 ** self.t)
        self.p.addcdiv_(self.m, denom, value=-step_size)


@pytest.mark.parametrize("max_norm", [1.0, None])
def test_fused_adamw_three_steps_match_hf_adamw(cuda, max_norm):
    from mmtg_b200.optim import FusedAdamW
    model, batch = _small_model(cuda)
    opt = FusedAdamW(model, lr=1e-3, max_grad_norm=max_norm)
    model.fused_train_step(batch, 3, 0.2)  # creates the flat buffers and real gradients
    P, W16, G = model._flat
    ref = _RefAdamW(P, 1e-3, max_norm=max_norm)
    for t in range(3):
        if t:
            opt.zero_grad()
            model.fused_train_step(batch, 3, 0.2)
        g = G.clone()
        if max_norm is not None and t == 0:
            # make sure the clip is ACTIVE whatever this small model's gradient norm is
            max_norm = 0.5 * torch.linalg.vector_norm(g).item()
            opt.max_grad_norm = ref.max_norm = max_norm
        opt.step()
        ref.step(g)
        # Adam's update is ~lr = 1e-3 per element; agreement to 2e-7 absolute (one fp32 ulp of the
        # O(1) LayerNorm weights) = 2e-4 of the update
        err = (P - ref.p).abs().max().item()
        assert err <= 2e-7, (t, err)
        assert (P - ref.p).abs().mean().item() <= 2e-9
        assert torch.allclose(opt._m, ref.m, rtol=1e-5, atol=1e-9)
        assert torch.allclose(opt._v, ref.v, rtol=1e-5, atol=1e-12)
        assert torch.equal(W16, P.to(torch.bfloat16)), "bf16 shadow != bf16(master) after the fused step"
        if max_norm is not None:
            assert abs(opt.grad_norm().item() - torch.linalg.vector_norm(g.double()).item()) <= 1e-3 * opt.grad_norm().item()
    assert int(opt._step_dev.item()) == 3


def test_fused_adamw_graph_replay_uses_device_step_and_lr(cuda):
    """The captured optimizer step reads the step counter and learning rate from device memory:
    replays advance the bias correction, and an LR scheduler's param_groups['lr'] reaches the
    kernel through sync_lr() (GraphedTrainStep calls it before every replay)."""
    from mmtg_b200.optim import FusedAdamW
    model, batch = _small_model(cuda)
    opt = FusedAdamW(model, lr=1e-3, max_grad_norm=1.0)
    model.fused_train_step(batch, 3, 0.2)
    P, W16, G = model._flat
    g = G.clone()
    ref = _RefAdamW(P, 1e-3, max_norm=1.0)
    opt.step()        # eager warm-up step (t = 1)
    ref.step(g)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            opt.step()
    torch.cuda.current_stream().wait_stream(side)
    # capture does not execute: counters unchanged
    assert int(opt._step_dev.item()) == 1
    for lr in (1e-3, 5e-4, 2.5e-4):
        opt.param_groups[0]["lr"] = lr  # what get_linear_schedule_with_warmup does
        opt.sync_lr()
        graph.replay()
        ref.lr = lr
        ref.step(g)
        torch.cuda.synchronize()
        assert (P - ref.p).abs().max().item() <= 2e-7
        assert torch.equal(W16, P.to(torch.bfloat16))
    assert int(opt._step_dev.item()) == 4


def test_fused_adamw_state_dict_roundtrip(cuda):
    from mmtg_b200.optim import FusedAdamW
    model, batch = _small_model(cuda)
    opt = FusedAdamW(model, lr=1e-3, max_grad_norm=1.0)
    model.fused_train_step(batch, 3, 0.2)
    opt.step()
    opt.step()
    sd = opt.state_dict()
    assert sd["mmtg_flat"]["step"] == 2
    opt2 = FusedAdamW(model, lr=1e-3, max_grad_norm=1.0)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2._m, opt._m) and torch.equal(opt2._v, opt._v) and int(opt2._step_dev.item()) == 2
    P = model._flat[0]
    before = P.clone()
    opt.step()
    after_a = P.clone()
    P.copy_(before)
    opt2.step()
    assert torch.equal(P, after_a), "resumed optimizer takes a different step"


def test_data_style_optimizer_updates_reach_the_forward(cuda):
    """transformers-4.12.3 AdamW updates weights through `p.data.add_()`, which
    does not bump tensor version counters. The bf16 weight shadow must still follow."""
    model, batch = _small_model(cuda)
    with torch.no_grad():
        _, _, a = model(batch)
        a = a.clone()
        for p in model.parameters():
            p.data.add_(0.01 * torch.sign(p.data))  # `.data`-style in-place update
        _, _, b = model(batch)
    assert (a - b).abs().max().item() > 1e-2, "forward ignored a .data-style weight update (stale bf16 shadow)"


def test_logits_are_released_by_refcount(cuda):
    """no reference cycle keeps the [B, L, V] logits alive after the caller
    drops them (checked with the cyclic GC disabled)."""
    import gc
    model, batch = _small_model(cuda)
    hf, kl, logits = model(batch)  # warm-up: workspace, flat gradient buffer, anchors
    (hf + kl).backward()
    del hf, kl, logits
    model.zero_grad(set_to_none=True)
    gc.collect()
    gc.disable()
    try:
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        for _ in range(3):
            hf, kl, logits = model(batch)
            (hf + kl).backward()
            n = logits.numel() * 4
            del hf, kl, logits
            model.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        grown = torch.cuda.memory_allocated() - base
        assert grown < n, f"{grown} bytes still allocated after dropping the outputs (logits = {n} bytes)"
    finally:
        gc.enable()
