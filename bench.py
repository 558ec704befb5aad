#!/usr/bin/env python
"""bench.py — MMTG hot-path benchmark (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload train|decode]
                  [--max-sent-length 20|40|60|98] [--batch B] [--dropout P]

One "step" = one pass of the training hot path over one synthetic batch of 32 samples per GPU:
MMTG.forward + MyLoss + 0.2*KL, backward (with the bucketed NCCL gradient all-reduce when N > 1),
global-norm clip and the AdamW update (restated src/train.py:188-197). Metric: train samples/s
(BASELINE.json configs[1]; weak scaling: the per-GPU batch is fixed at 32).

  value : device-timed (CUDA events, max over ranks) with the batch already resident in HBM
  e2e   : the same step through the public Python surface with the batch in PINNED HOST memory,
          host->device copies and the device->host loss read inside the timed region
  roofline     : tensor bound for the dominant kernel (the tcgen05 GEMM), from per-launch CUDA
                 events recorded by the library in extra, separately run, profiled steps
  cpu_baseline : the CPU oracle's train step on the host cores (bounded sample)

The default N = 1 run ALSO measures BASELINE.json configs[3] (KV-cached generation, batch 64, 220
positions) and nests it in the same JSON line under "decode" (metric decode tokens/s, HBM roofline,
its own e2e and cpu_baseline); `--workload decode` prints that line alone. `--max-sent-length` > 20
and `--stage` / `--neg-frac` are the extended-length / negative-ratio runs of configs[4].
`torch_eager_gpu_baseline` (N = 1): the unmodified reference run by PyTorch eager on the same GPU and
HF GPT-2 (SDPA, bf16 autocast, AdamW) on the same shapes — the library kernels to beat.

`--impl reference` / `cpu_baseline` time the reference's own implementation on the host cores:
the UNMODIFIED reference staged under oracle/_ref (kind "reference", oracle/build_ref.py) when it is
there, else the oracle port (kind "port"). Protocol: 2 warm-ups, median of >= 5.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

TRAIN_GFLOP_PER_SAMPLE = 140.2  # BASELINE.md §3, L = 236
PER_GPU_BATCH = 32
ALPHA = 0.2
STAGE = 3


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region: NVML polled every 50 ms (nvidia-smi
    subprocesses are too slow when 8 ranks sample at once), nvidia-smi as the fallback."""
    _REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.rows.append((float(sm), float(mx), {name for name, bit in self._REASONS if mask & bit}))
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout
        c = [x.strip() for x in out.strip().split(",")]
        if len(c) >= 6 and c[0].replace(".", "").isdigit():
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.rows.append((float(c[0]), float(c[1]), {nm for nm, v in zip(names, c[2:6]) if v.lower().startswith("active")}))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop_evt.wait(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        self._stop_evt.set()
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = set()
        for r in self.rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


class _Tok:
    """The only tokenizer calls on the reference's sampling path (src/generate.py:133-136)."""
    _ids = {"[PAD]": 0, "[#START#]": 1, "[#EOS#]": 2, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102}

    def convert_tokens_to_ids(self, t):
        return self._ids[t]


_REF_CACHE = {}


def _reference(device="cpu"):
    """(kind, objects): the unmodified reference (oracle/_ref or /root/reference) or the oracle port."""
    key = str(device)
    if key in _REF_CACHE:
        return _REF_CACHE[key]
    from mmtg_b200 import synth
    from oracle import ref_import
    table = synth.make_token_table()
    sd = synth.make_state_dict(0)
    if ref_import.available():
        model, crit, gen, dc, cfgs = ref_import.load_reference(table, sd)
        model.to(device)
        out = ("reference", dict(model=model, crit=crit, gen=gen, dc=dc, table=table, sd=sd))
    else:
        out = ("port", dict(table=table, sd=sd))
    _REF_CACHE[key] = out
    return out


def reference_train_step_times(batch_size, reps, warmups, threads, device="cpu"):
    """The reference's train step (src/train.py:188-197: forward, MyLoss + alpha*KL, backward,
    clip_grad_norm_(1.0), AdamW(lr 1e-5, eps 1e-6), zero_grad) in fp32. Returns (kind, [s/step])."""
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config
    torch.set_num_threads(threads)
    kind, R = _reference(device)
    batch = {k: v.to(device) for k, v in synth.batch_to_torch(synth.make_batch(batch_size, seed=1234)).items()}
    on_gpu = str(device).startswith("cuda")
    if kind == "reference":
        model, crit = R["model"], R["crit"]
        model.train()  # train.py never calls eval() inside the loop: GPT-2 dropout is live
        model.train_flag = True
        opt = torch.optim.AdamW(model.parameters(), lr=1e-5, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0)

        def step():
            _l, kl, logits = model.forward(batch)
            loss = crit(logits.contiguous(), batch["targets"], batch["rating"], STAGE)
            total = loss.mean() + ALPHA * kl.mean()
            total.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            opt.zero_grad()
    else:
        from oracle import mmtg_oracle as O
        table = torch.from_numpy(R["table"]).to(device)
        params = {k: v.to(device).clone().requires_grad_(True) for k, v in R["sd"].items() if k != "decoder.gpt2.lm_head.weight"}
        opt = torch.optim.AdamW(list(params.values()), lr=1e-5, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0)
        params["decoder.gpt2.lm_head.weight"] = params["decoder.gpt2.transformer.wte.weight"]

        def step():
            opt.zero_grad(set_to_none=True)
            hf, kl, logits = O.mmtg_forward(params, table, batch, data_config(), True)
            total = O.my_loss(logits, batch["targets"], batch["rating"], STAGE).mean() + ALPHA * kl.mean()
            total.backward()
            torch.nn.utils.clip_grad_norm_(opt.param_groups[0]["params"], 1.0)
            opt.step()
    ts = []
    for i in range(warmups + reps):
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        if on_gpu:
            torch.cuda.synchronize()
        if i >= warmups:
            ts.append(time.perf_counter() - t0)
    return kind, ts


def reference_decode_times(positions, reps, warmups, threads):
    """The reference's sample_sequence (src/generate.py:97-145: batch 1, greedy, full-prefix
    recompute per token, no KV cache) on the host cores. Returns (kind, [s/call])."""
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config
    torch.set_num_threads(threads)
    kind, R = _reference("cpu")
    one = synth.make_batch(1, seed=1234)
    start = {k: v[0] for k, v in one.items() if k != "rating"}
    start["targets"] = np.asarray([1])
    ts = []
    for i in range(warmups + reps):
        t0 = time.perf_counter()
        if kind == "reference":
            R["model"].eval()
            R["model"].train_flag = False
            R["gen"].sample_sequence(R["model"], dict(start), positions, _Tok(), temperature=1.0, top_k=1, top_p=0.0,
                                     repitition_penalty=1.0, device="cpu")
        else:
            from oracle import mmtg_oracle as O
            O.sample_sequence(R["sd"], torch.from_numpy(R["table"]), start, positions, data_config(), temperature=1.0,
                              top_k=1, top_p=0.0, repitition_penalty=1.0)
        if i >= warmups:
            ts.append(time.perf_counter() - t0)
    return kind, ts


def hf_gpt2_sdpa_step_time(batch_size, L, reps, dev):
    """Strongest library baseline for the decoder: transformers GPT2LMHeadModel (SDPA attention) of
    src/config/model_config.json on [B, L, 768] inputs_embeds under bf16 autocast, forward + HF CE
    loss + backward + torch AdamW (fused) — cuBLASLt / flash-SDPA sm_100 kernels, optimizer included.
    It omits the encoder side, the embedding build and MyLoss, so it flatters the baseline."""
    import transformers
    cfg = transformers.GPT2Config(vocab_size=13317, n_positions=1024, n_embd=768, n_layer=12, n_head=12)
    m = transformers.GPT2LMHeadModel(cfg).to(dev).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-5, eps=1e-6, weight_decay=0.0, fused=True)
    x = torch.randn(batch_size, L, 768, device=dev)
    labels = torch.randint(0, 13317, (batch_size, L), device=dev)
    mask = torch.ones(batch_size, L, dtype=torch.long, device=dev)
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(inputs_embeds=x, attention_mask=mask, labels=labels)
        out.loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) / 1e3)
    del m, opt, out
    torch.cuda.empty_cache()
    return float(np.median(ts))


def gpu_eager_step_time(batch_size, reps, dev, autocast):
    """The same restated reference step run by PyTorch eager (cuBLAS / SDPA-free plain ops) on the
    B200 itself: "the Blackwell kernels to beat" (SURVEY §8d second reported baseline).
    Oracle code, baseline leg only. Returns s/step (CUDA events)."""
    from mmtg_b200 import synth
    from mmtg_b200.configs import data_config
    from oracle import mmtg_oracle as O
    table = torch.from_numpy(synth.make_token_table()).to(dev)
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(0).items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "decoder.gpt2.lm_head.weight"}
    params["decoder.gpt2.lm_head.weight"] = params["decoder.gpt2.transformer.wte.weight"]
    batch = {k: v.to(dev) for k, v in synth.batch_to_torch(synth.make_batch(batch_size, seed=1234)).items()}
    ts = []
    for i in range(reps + 1):
        for p in params.values():
            p.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            hf, kl, logits = O.mmtg_forward(params, table, batch, data_config(), True)
        total = O.my_loss(logits.float(), batch["targets"], batch["rating"], STAGE).mean() + ALPHA * kl.float().mean()
        total.backward()
        e1.record()
        torch.cuda.synchronize()
        if i > 0:
            ts.append(e0.elapsed_time(e1) / 1e3)
    del params, sd, logits, total
    torch.cuda.empty_cache()
    return float(np.median(ts))


def measure_decode(args, dev, cpu_baseline=True):
    """BASELINE.json configs[3]: greedy / top-k KV-cached generation, batch 64, 220 positions, 1 GPU
    (generation does not shard: replicas only). One step = one whole generation call through the
    public surface (host arrays in, token lists out). Returns the result dict."""
    from mmtg_b200 import _lib, synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200 import generate as G
    from mmtg_b200.generate import sample_sequence_batch
    from mmtg_b200.model import MMTG
    B, LENGTH = 64, 220
    model = MMTG(model_cfgs, data_config(), 13317, train_flag=False, token_table=synth.make_token_table())
    model.load_state_dict(synth.make_state_dict(0))
    model.to(dev)
    batch = synth.make_batch(B, seed=1234)
    starts = {k: v for k, v in batch.items() if k != "rating"}
    starts["targets"] = np.ones((B, 1), np.int64)
    presets = {"greedy": dict(temperature=1.0, top_k=1, top_p=0.0, repitition_penalty=1.0),
               "topk10_p0.7": dict(temperature=1.1, top_k=10, top_p=0.7, repitition_penalty=1.5)}
    W, K = max(args.warmup, 3), max(1, min(args.steps, 20))
    sampler = ClockSampler(dev.index or 0)
    sampler.start()

    def timed(kw):
        for _ in range(W):
            sample_sequence_batch(model, starts, LENGTH, device=str(dev), **kw)
        torch.cuda.synchronize()
        l0 = _lib.launch_count() + G.replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(K):
            sample_sequence_batch(model, starts, LENGTH, device=str(dev), seed=i, **kw)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / K
        return wall, (_lib.launch_count() + G.replayed_launches - l0) // K, e0.elapsed_time(e1) / 1e3 / K

    res = {name: timed(kw) for name, kw in presets.items()}
    # device-only time of the decode loop itself (positions 1..219, CUDA events inside the call)
    G.last_steps_ms = None
    os.environ["MMTG_GEN_EVENTS"] = "1"
    try:
        sample_sequence_batch(model, starts, LENGTH, device=str(dev), **presets["greedy"])
        steps_ms = G.last_steps_ms
    finally:
        os.environ.pop("MMTG_GEN_EVENTS", None)
    clocks = sampler.summary()
    pk, pk_src = peaks()
    sec, launches, dev_sec = res["greedy"]
    # algorithmic bytes: 193.2 MB of bf16 weights per position + 36,864 B of K/V per cached key and row
    gbytes = (LENGTH * 193.2e6 + sum(36864.0 * (15 + j) for j in range(LENGTH)) * B) / 1e9
    h2d = sum(np.asarray(v).nbytes for v in starts.values())
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_decode_traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]  # one launch at a late position (ncu --set full)
    except Exception:
        pass
    line = {
        "metric": "decode tokens/s", "value": B * LENGTH / sec, "unit": "tokens/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "MMTG KV-cached greedy generation, batch 64, 220 positions, 1 GPU (BASELINE.json configs[3]); "
                               "one step = one sample_sequence_batch call (prefill + 219 decoded positions, CUDA-graph replay)",
                   "batch": B, "length": LENGTH, "l2": "per-position working set (193 MB weights + KV) exceeds L2"},
        "e2e": {"value": B * LENGTH / sec, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": B * (LENGTH + 1) * 4,
                "note": "the public call takes host arrays and returns host token lists: value == e2e (wall clock around the call)"},
        "device_ms_per_step": dev_sec * 1e3,
        "decode_loop_ms": steps_ms, "us_per_position": (steps_ms * 1e3 / (LENGTH - 1)) if steps_ms else None,
        "gpu_launches": launches * K, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": gbytes / sec, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": gbytes / sec / pk["hbm_gbs"], "traffic": traffic,
                     "x_of_floor": sec / (gbytes / pk["hbm_gbs"]),
                     "kernel": "decode_mega_kernel (ONE persistent launch for all decoded positions of a call: embedding, projector, 12 blocks, lm_head and sampler in-kernel, grid barriers between phases); algorithmic bytes = bf16 weights + K/V per position, whole call",
                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({pk_src})"},
        "topk_preset_tokens_per_s": B * LENGTH / res["topk10_p0.7"][0],
    }
    if cpu_baseline:
        threads = os.cpu_count() or 1
        kind, ts = reference_decode_times(50, 3, 1, threads)
        med = float(np.median(ts))
        line["cpu_baseline"] = {"value": 50 / med, "unit": "tokens/s", "cores": threads, "kind": kind,
                                "sample": f"median of {len(ts)} calls after 1 warm-up of sample_sequence greedy, 50 positions, batch 1, "
                                          "full-prefix recompute per token (src/generate.py:97-145), fp32"}
    return line


def run_decode(args, rank, local_rank):
    if rank != 0:
        return
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    print(json.dumps(measure_decode(args, dev, cpu_baseline=not args.no_cpu_baseline)), flush=True)


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation on the host cores, same metric/config."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    bs = 2
    W, K = max(1, min(args.warmup, 3)), max(1, min(args.steps, 20))
    if args.workload == "decode":
        kind, ts = reference_decode_times(50, max(1, min(K, 5)), 1, threads)
        sec = float(np.median(ts))
        val = 50 / sec
        line = {
            "impl": "reference", "metric": "decode tokens/s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": len(ts), "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MMTG greedy generation (BASELINE.json configs[3]) -- reference sample_sequence on the host cores: "
                                   "batch 1, no KV cache, bounded sample of 50 positions per step"},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": threads, "kind": kind,
                             "sample": f"median of {len(ts)} calls of 50 positions after 1 warm-up"},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    kind, ts = reference_train_step_times(bs, K, W, threads)
    sec = float(np.mean(ts))  # K timed steps back to back, like the GPU arm
    val = bs / sec
    what = ("the UNMODIFIED reference (oracle/_ref: src/model.py + src/loss.py, torch AdamW eps 1e-6, train mode)"
            if kind == "reference" else "oracle port (oracle/mmtg_oracle.py)")
    line = {
        "impl": "reference", "metric": "train samples/s", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": len(ts), "warmup": W, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "MMTG train step (fwd + MyLoss stage 3 + 0.2*KL, bwd, clip 1.0, AdamW), L=236, V=13317 (BASELINE.json configs[1]) -- " + what + " on the host cores, fp32, bounded sample: batch 2 per step"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": kind,
                         "sample": f"{len(ts)} timed steps of batch {bs} (L=236) after {W} warm-ups; median {bs / float(np.median(ts)):.2f} samples/s"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mmtg_b200")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="train", choices=["train", "decode"],
                    help="train: BASELINE.json configs[1]/[2]/[4]; decode: configs[3] (KV-cached generation, 1 GPU)")
    ap.add_argument("--max-sent-length", type=int, default=20,
                    help="tokens per lyric sentence (20 = the reference's L = 236; configs[4] sweeps 40/60/98)")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="GPT-2 embd/resid/attn dropout of the training forward (reference default 0.1)")
    ap.add_argument("--stage", type=int, default=STAGE, choices=[1, 2, 3],
                    help="curriculum stage of MyLoss (src/loss.py:57-60); stages 1/2 also filter rows by rating "
                         "(src/train.py:178-183), which gives ragged per-rank batches")
    ap.add_argument("--neg-frac", type=float, default=None,
                    help="configs[4] negative-ratio sweep: force this fraction of each rank's rows to be negatives "
                         "(rating 1-2; the rest 4-5) instead of uniform ratings")
    ap.add_argument("--no-decode", action="store_true", help="skip the nested configs[3] decode measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "decode":
        run_decode(args, rank, local_rank)
        return

    import torch.distributed as dist
    from mmtg_b200 import _lib, synth
    from mmtg_b200.configs import data_config, model_cfgs
    from mmtg_b200.loss import MyLoss
    from mmtg_b200.model import MMTG
    from mmtg_b200.optim import FusedAdamW
    from mmtg_b200.parallel import GradSync

    W = max(args.warmup, 3)
    K = args.steps
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    B = args.batch
    dcfg = data_config(max_sent_length=args.max_sent_length)
    L = dcfg["topic_prompt_length"] + dcfg["max_seq_length"] + 1
    gflop_per_sample = {236: 140.2, 436: 263.6, 636: 391.4, 1016: 646.4}.get(L, 140.2 * L / 236.0)
    table = synth.make_token_table()
    model = MMTG(model_cfgs, dcfg, 13317, train_flag=True, token_table=table)
    model.load_state_dict(synth.make_state_dict(0))  # identical replicas on every rank
    model.set_dropout(args.dropout, args.dropout, args.dropout)
    model.to(dev)
    equal_scale = 1.0
    if world > 1:
        # default: AVG all-reduce -> NCCL picks RING / LL128 on 32 channels. MMTG_DDP_SUM=1: SUM all-reduce of
        # gradients whose loss was pre-scaled by 1 / world -> NCCL picks NVLS / SIMPLE on 24 channels (in-switch
        # reduction). Measured at 8 GPUs (profiles/r2_nccl_n8_algo.txt): 9.56 ms/step vs 9.65 ms/step — the NVLS
        # kernels slow the concurrent backward GEMMs a little more (55.3 vs 52.6 us per dominant launch).
        if os.environ.get("MMTG_DDP_SUM") == "1":
            model.grad_sync = GradSync(average=False)
            equal_scale = 1.0 / world
        else:
            model.grad_sync = GradSync()
    crit = MyLoss(dcfg, model_cfgs)
    opt = FusedAdamW(model, lr=1e-5, max_grad_norm=1.0)
    stage = args.stage
    ratings = None
    if args.neg_frac is not None:  # configs[4]: fixed negative fraction (ratings 1-2 vs 4-5; 3 never drawn)
        rr = np.random.default_rng(99 + rank)
        n_neg = int(round(args.neg_frac * B))
        ratings = np.concatenate([rr.integers(1, 3, n_neg), rr.integers(4, 6, B - n_neg)])
        rr.shuffle(ratings)
    host = synth.batch_to_torch(synth.make_batch(B, seed=1234 + rank, data_config=dcfg, ratings=ratings))
    b_before = B
    grad_scale = equal_scale
    if stage in (1, 2):  # the reference's rating filter (src/train.py:178-183): per-rank row counts differ
        from mmtg_b200.curriculum import stage_row_indices
        from mmtg_b200.parallel import ragged_batch_scale
        idx = stage_row_indices(host["rating"], stage)
        host = {k: v[idx] for k, v in host.items()}
        B = int(len(idx))
        assert B > 0, "no row of this rank survives the curriculum filter"
        if world > 1:
            model.grad_sync = GradSync(average=False)  # SUM of gradients scaled by B_local / B_global
            grad_scale = ragged_batch_scale(B, device=dev)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    # L2 flush buffer (> 126 MB); the step's own working set (~3 GB of activations) already exceeds L2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def eager_step(batch):
        hf, kl, logits = model(batch)
        loss = crit(logits, batch["targets"], batch["rating"], stage)
        total = (loss.mean() + ALPHA * kl.mean()) * grad_scale
        total.backward()
        opt.step()
        opt.zero_grad()
        return total

    # ---- parity gate: the first step (dropout off) against the committed CPU-oracle golden ----
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", "bench_b32_step.json")
    if rank == 0 and B == 32 and L == 236 and stage == 3 and args.neg_frac is None and os.path.isfile(gpath):
        with open(gpath) as f:
            gold = json.load(f)
        model.set_dropout(0.0, 0.0, 0.0)
        sync_save, model.grad_sync = model.grad_sync, None  # rank-0-only check: no collective
        t0_, l0_, k0_ = model.fused_train_step(resident, 3, ALPHA)
        torch.cuda.synchronize()
        model.grad_sync = sync_save
        model._flat[2].zero_()
        model.set_dropout(args.dropout, args.dropout, args.dropout)
        parity = {"myloss": float(l0_), "myloss_oracle": gold["myloss_stage3"], "kl": float(k0_), "kl_oracle": gold["kl"],
                  "tolerance": "|d myloss| <= 2e-3, |d kl| <= 5e-3 (BASELINE.md §5)", "source": "tests/golden/bench_b32_step.json"}
        parity["ok"] = bool(abs(parity["myloss"] - gold["myloss_stage3"]) <= 2e-3 * max(1.0, abs(gold["myloss_stage3"]))
                            and abs(parity["kl"] - gold["kl"]) <= 5e-3)
        if not parity["ok"]:
            raise SystemExit(f"bench.py parity gate FAILED: {parity}")

    # fixed shapes -> the whole step (incl. the bucketed NCCL all-reduces on the side stream) is
    # replayed from one CUDA graph (launch-bound otherwise); MMTG_GRAPH=0 = eager launches.
    use_graph = os.environ.get("MMTG_GRAPH", "1") == "1"
    if use_graph:
        from mmtg_b200.graph import GraphedTrainStep
        step = GraphedTrainStep(model, crit, opt, resident, alpha=ALPHA, stage=stage, warmup=3, grad_scale=grad_scale)
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(resident)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-timed, inputs resident ----
    l0 = lib.mmtg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.profiler.start()  # ncu --profile-from-start off: capture exactly the timed steps
    e0.record()
    for _ in range(K):
        step(resident)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = int(lib.mmtg_launch_count() - l0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    # ---- end to end through the public API: pinned host batch -> H2D -> step -> loss.item() ----
    def e2e_step():
        if use_graph:  # pinned host -> device staging (copy stream, one step ahead) -> static buffers -> replay
            out = step(pinned)
            step.prefetch(pinned)  # the next step's inputs cross PCIe while this step computes
            return out
        return step({k: v.to(dev, non_blocking=True) for k, v in pinned.items()})

    # every step's loss is read back on the host, two steps late (asynchronous D2H into a pinned
    # ring), the way a training loop logs it without stalling the launch queue
    ring = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    ring_evt = [torch.cuda.Event(), torch.cuda.Event()]
    losses = []

    def run_e2e(n):
        for k in range(n):
            total = e2e_step()
            slot = k % 2
            if k >= 2:
                ring_evt[slot].synchronize()
                losses.append(float(ring[slot][0]))
            ring[slot].copy_(total.detach().reshape(1), non_blocking=True)
            ring_evt[slot].record()
        for k in range(max(0, n - 2), n):
            ring_evt[k % 2].synchronize()
            losses.append(float(ring[k % 2][0]))

    if use_graph:
        step.prefetch(pinned)
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    losses.clear()
    run_e2e(K)
    e3.record()
    barrier()
    assert len(losses) == K and all(np.isfinite(losses))
    e2e_wall = time.perf_counter() - t0
    e2e_ms = torch.tensor([max(e2.elapsed_time(e3), e2e_wall * 1e3)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.summary()

    # ---- roofline: per-launch CUDA events of the dominant kernel (tcgen05 GEMM), 2 extra steps ----
    import ctypes as C
    lib.mmtg_prof_reset()
    lib.mmtg_prof_enable(1)
    lib.mmtg_set_wgrad_side_stream(0)  # per-launch event times must not include a concurrent kernel
    PROF_STEPS = 2
    lp0 = lib.mmtg_launch_count()
    for _ in range(PROF_STEPS):  # profiled steps always launch eagerly (events per launch)
        flush.zero_()
        eager_step(resident)
    torch.cuda.synchronize()
    lib.mmtg_prof_enable(0)
    lib.mmtg_set_wgrad_side_stream(1)
    launches_per_step = int(lib.mmtg_launch_count() - lp0) // PROF_STEPS
    if use_graph:  # replays bypass the library's launch counter: same kernels, K replays
        launches = launches_per_step * K
    classes = {}
    for cls, name in ((0, "gemm_tcgen05"), (1, "attention"), (2, "row_kernels")):
        t, f, b, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        lib.mmtg_prof_collect(cls, C.byref(t), C.byref(f), C.byref(b), C.byref(n))
        classes[name] = {"ms_per_step": t.value / PROF_STEPS, "launches_per_step": n.value // PROF_STEPS,
                         "gflop_per_step": f.value / PROF_STEPS / 1e9, "gbytes_per_step": b.value / PROF_STEPS / 1e9,
                         "achieved_gbs": b.value / max(t.value, 1e-9) / 1e6, "achieved_tflops": f.value / max(t.value, 1e-9) / 1e9}
    dump = os.environ.get("MMTG_PROF_DUMP", f"/tmp/mmtg_prof_{os.getpid()}.csv")
    lib.mmtg_prof_dump(dump.encode())
    lib.mmtg_prof_reset()
    pk, pk_src = peaks()
    gemm = classes["gemm_tcgen05"]
    all_gemm_tflops = gemm["gflop_per_step"] / max(gemm["ms_per_step"], 1e-9)  # GFLOP/ms = TFLOP/s
    # Which measured peak applies: the timed region is a fraction of a second and the profiled steps
    # ~10 ms each, far from the 4 s soak behind bf16_tflops_sustained (taken at a median SM clock of
    # 1342 MHz). When the clock sampled DURING the timed region stays >= 0.95 of max, the GPU is in
    # the burst regime and the burst peak is the honest denominator.
    burst = bool(clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
    peak = pk["bf16_tflops"] if burst else pk["bf16_tflops_sustained"]
    peak_name = "bf16_tflops (burst: sampled SM clock >= 0.95 of max)" if burst else "bf16_tflops_sustained (SM clock below 0.95 of max in the timed region)"
    # dominant kernel: the tcgen05 GEMM on the 7552x3072x768-FLOP shape family (c_fc forward, its
    # dgrad pair and the two 768x3072 wgrads: 84 of the 186 GEMM launches, ~55 % of GEMM time)
    dom_flops = 2.0 * (B * L) * 3072 * 768
    dom_ms, dom_n = 0.0, 0
    try:
        import csv as _csv
        for row in _csv.DictReader(open(dump)):
            if row["class"] == "0" and abs(float(row["flops"]) - dom_flops) < 1.0:
                dom_ms += float(row["ms"])
                dom_n += 1
    except Exception:
        pass
    achieved = dom_flops * dom_n / max(dom_ms, 1e-9) / 1e9 if dom_n else all_gemm_tflops
    traffic = None
    for name in ("r2_gemm_dominant_traffic.json", "r1_gemm_dominant_traffic.json"):
        try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (latest round first)
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f)["dram_bytes_per_launch"]
            break
        except Exception:
            pass

    # reported context for the tensor roofline: the dominant contraction (M x 3072 x 768, plain bf16
    # output) launched alone with a cold L2 - this library next to cuBLAS (torch.matmul). The burst
    # peak in MEASURED_PEAKS.json is cuBLAS at 8192^3; at K = 768 a tile is 12 k-blocks long and
    # neither kernel gets close to it, so this is the like-for-like reference of `frac`.
    same_shape = None
    if world == 1:
        try:
            from mmtg_b200 import ops as _ops
            Mg = B * L
            ga = torch.randn(Mg, 768, device=dev).to(torch.bfloat16)
            gw = (torch.randn(3072, 768, device=dev) * 0.05).to(torch.bfloat16)
            go = torch.empty(Mg, 3072, device=dev, dtype=torch.bfloat16)

            def _med(fn, n=7):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(n):
                    flush.zero_()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    fn()
                    a1.record()
                    torch.cuda.synchronize()
                    ts.append(a0.elapsed_time(a1))
                return sorted(ts)[len(ts) // 2]
            t_own = _med(lambda: _ops.gemm(ga, gw, go, M=Mg, N=3072, K=768))
            t_lib = _med(lambda: torch.matmul(ga, gw.t(), out=go))
            same_shape = {"shape": [Mg, 3072, 768], "this_library_us": 1e3 * t_own, "this_library_tflops": dom_flops / t_own / 1e9,
                          "cublas_us": 1e3 * t_lib, "cublas_tflops": dom_flops / t_lib / 1e9,
                          "how": "one launch at a time, 256 MB L2 flush before each, CUDA events, median of 7; cuBLAS = torch.matmul (bf16)"}
            del ga, gw, go
        except Exception as exc:  # a reported baseline must never take the bench down
            same_shape = {"error": repr(exc)}

    gb = torch.tensor([float(B)], device=dev)
    if world > 1:
        dist.all_reduce(gb, op=dist.ReduceOp.SUM)  # ragged per-rank batches after a stage-1/2 filter
    global_batch = int(gb.item())
    value = global_batch * K / (ms_total / 1e3)
    e2e_value = global_batch * K / (e2e_ms.item() / 1e3)
    if rank == 0:
        line = {
            "metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "MMTG train step (fwd + MyLoss stage %d + 0.2*KL, bwd, clip 1.0, AdamW) bf16 GEMMs / fp32 master+residual, batch %d per GPU%s, L=%d, V=13317 (BASELINE.json configs[%d]); GPT-2 embd/resid/attn dropout p=%g (fused counter-based masks)" % (stage, b_before, "" if B == b_before else " (%d rows on rank 0 after the stage-%d rating filter)" % (B, stage), L, 1 if (L == 236 and stage == 3 and args.neg_frac is None) else 4, args.dropout),
                       "stage": stage, "neg_frac": args.neg_frac,
                       "global_batch": global_batch, "seq_len": L,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "launch": "cuda_graph" if use_graph else "eager",
                       "l2": "working set (~3 GB activations/step) exceeds L2; 256 MB flush before profiled steps"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms.item() / K},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": f"gemm_bf16_tcgen05_kernel, 2*{B * L}*3072*768 FLOP per launch (c_fc fwd, its dgrads, the 768x3072 wgrads); CUDA events per launch in 2 eagerly launched steps",
                         "launches_per_step": dom_n // PROF_STEPS if dom_n else None,
                         "avg_launch_us": 1e3 * dom_ms / dom_n if dom_n else None,
                         "peak_source": f"MEASURED_PEAKS.json {peak_name} ({pk_src})",
                         "frac_of_sustained_peak": achieved / pk["bf16_tflops_sustained"],
                         "all_gemm_launches_tflops": all_gemm_tflops,
                         "step_model_flops_frac": (value / world) * gflop_per_sample / 1e3 / peak,
                         "same_shape_alone": same_shape},
            "breakdown": classes,
        }
        if parity is not None:
            line["parity_check"] = parity
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            kind, ts2 = reference_train_step_times(2, 5, 2, threads)
            _k, ts8 = reference_train_step_times(8, 5, 2, threads)
            line["cpu_baseline"] = {"value": 2 / float(np.median(ts2)), "unit": "samples/s", "cores": threads, "kind": kind,
                                    "batch8_samples_per_s": 8 / float(np.median(ts8)),
                                    "sample": "median of 5 timed train steps after 2 warm-ups (BASELINE.md §4) at batch 2 (value) and batch 8, L=236: "
                                              + ("the unmodified reference (oracle/_ref: model.py forward + loss.py MyLoss) " if kind == "reference" else "oracle/mmtg_oracle.py ")
                                              + "+ 0.2*KL + autograd bwd + clip 1.0 + torch AdamW(eps 1e-6), fp32, all host threads"}
            try:  # reported baselines on this GPU: library kernels under PyTorch eager
                bs = 32
                eg = {"unit": "samples/s", "batch": bs}
                kind_g, tsg = reference_train_step_times(bs, 3, 1, threads, device=dev)
                eg["reference_fp32" if kind_g == "reference" else "oracle_port_fp32"] = bs / float(np.median(tsg))
                _REF_CACHE.clear()
                torch.cuda.empty_cache()
                eg["oracle_port_bf16_autocast_fwd_bwd"] = bs / gpu_eager_step_time(bs, 3, dev, True)
                eg["hf_gpt2_sdpa_bf16_adamw"] = bs / hf_gpt2_sdpa_step_time(bs, L, 5, dev)
                eg["sample"] = ("reference_fp32: the unmodified reference's train step on this GPU (its per-token Python embedding loop included); "
                                "oracle_port_bf16_autocast_fwd_bwd: vectorised port, math attention, no optimizer; "
                                "hf_gpt2_sdpa_bf16_adamw: transformers GPT2LMHeadModel (SDPA) decoder only + fused AdamW; median of 3-5 steps, CUDA events")
                line["torch_eager_gpu_baseline"] = eg
            except Exception as e:  # reported baseline only: never fails the bench
                line["torch_eager_gpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_decode:
            del step
            model.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            line["decode"] = measure_decode(args, dev, cpu_baseline=not args.no_cpu_baseline)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
