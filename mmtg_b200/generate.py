"""Drop-in generation surface (mirrors /root/reference/src/generate.py:64-145).

`sample_sequence(model, start_input, length, tokenizer, temperature, top_k, top_p,
repitition_penalty, device)` keeps the reference's signature (sic: `repitition`) and return value
(the token list WITHOUT the last appended token). Underneath, the reference's "re-run the whole
model on the whole prefix per token" loop is replaced by a KV-cached decode engine:
one training-style forward over [prompt | first token], then ONE persistent-kernel launch that
decodes every remaining position (mmtg_decode_steps_fused: embedding, projector, 12 blocks, lm_head
and the sampler per position, grid barriers in between; B <= 64), or per-op launches replayed from
a CUDA graph (mmtg_decode_step + mmtg_sample_rows; B > 64). `sample_sequence_batch` runs B independent
batch-1 reference runs at once (BASELINE.json configs[3]: batch 64).
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch

from . import _lib

_FLOAT_KEYS = ("topic_emb", "img_embs", "r_embs")
_INT_KEYS = ("topic_ids", "tpw_attention_mask", "tpw_type_ids")


def top_k_top_p_filtering(logits, top_k=0, top_p=0.0, filter_value=-float("Inf")):
    """src/generate.py:64-94 on the device sampler: returns `logits` with every token the
    reference would filter set to `filter_value` (1-D logits; temperature 1, no penalty).
    top_k = 0 with top_p > 0 is the pure nucleus filter (the function's defaults filter nothing)."""
    assert logits.dim() == 1
    if not logits.is_cuda:
        raise _lib.MMTGError("mmtg_b200.top_k_top_p_filtering runs on CUDA only")
    V = logits.numel()
    probs = _filtered_distribution(logits.detach().float().view(1, V), top_k, top_p, ban=False)
    finite = torch.isfinite(logits)
    logits[(probs[0] <= 0) & finite] = filter_value
    return logits


def _filtered_distribution(logits2d, top_k, top_p, temperature=1.0, ban=True, history=None, rep_penalty=1.0):
    """Dense [B, V] probabilities of the sampler's filtered distribution (0 = filtered out).
    `history`: optional [B, n] int token ids the repetition penalty is applied over."""
    Bn, V = logits2d.shape
    dev = logits2d.device
    z = logits2d.contiguous()
    if history is None:
        gen = torch.full((Bn, 4), 5, dtype=torch.int32, device=dev)  # dummy history (penalty 1: unused)
        j = torch.zeros(1, dtype=torch.int32, device=dev)
    else:
        h = torch.as_tensor(history, dtype=torch.int32, device=dev).view(Bn, -1)
        gen = torch.cat([h, torch.zeros(Bn, 1, dtype=torch.int32, device=dev)], 1).contiguous()
        j = torch.full((1,), h.shape[1] - 1, dtype=torch.int32, device=dev)
    dbg = torch.empty(Bn, V, device=dev)
    _lib.check(_lib.lib().mmtg_sample_rows(C.c_void_p(z.data_ptr()), C.c_int64(z.stride(0)), C.c_void_p(gen.data_ptr()),
                                           gen.shape[1], C.c_void_p(j.data_ptr()), Bn, V, 1 << 20, C.c_float(temperature),
                                           int(top_k), C.c_float(top_p), C.c_float(rep_penalty), C.c_uint64(0), None, int(ban),
                                           C.c_void_p(dbg.data_ptr()), C.c_void_p(_lib.stream_ptr())),
               "mmtg_sample_rows")
    return dbg


def _collate(start_inputs):
    if isinstance(start_inputs, dict):
        return {k: np.asarray(v) for k, v in start_inputs.items()}
    return {k: np.stack([np.asarray(s[k]) for s in start_inputs]) for k in start_inputs[0] if k != "rating"}


def returned_length(length, sent_len):
    """Length of the list the reference returns: `generated` is captured at the last iteration
    that runs the model (src/generate.py:126,144), i.e. targets[: i_last + 1]."""
    i_last = -1
    for i in range(length):
        if i > 0 and (i + 2) % sent_len in (0, 1):
            continue
        i_last = i
    return i_last + 1


# kernels launched through CUDA-graph replays (they bypass the library's launch counter)
replayed_launches = 0
# MMTG_GEN_EVENTS=1: device time (ms, CUDA events) of the last call's decode loop (positions after the prefill)
last_steps_ms = None


class _DecodeSession:
    """Per-(batch, length, sampling preset) decode state kept on the model: KV-cache workspace,
    token / step-index / logits buffers and the captured CUDA graph of one decode step, so
    repeated generation calls pay neither allocation nor capture."""

    def __init__(self, model, d, length, Bn, dev, sampling):
        lib = _lib.lib()
        self.Lmax = d.P + length + 1
        self.gen_ld = length + 1
        self.dws = torch.empty(lib.mmtg_decode_workspace_bytes(C.byref(d), self.Lmax), dtype=torch.uint8, device=dev)
        self.gen = torch.zeros(Bn, self.gen_ld, dtype=torch.int32, device=dev)
        self.j = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seed = torch.zeros(1, dtype=torch.int64, device=dev)
        self.step_logits = torch.empty(Bn, d.V, device=dev)
        self.graph = None
        self.sampling = sampling


class _SectionTimer:
    """MMTG_GEN_TIMING=1: print synchronized wall time of each section of a generation call."""

    def __init__(self):
        self.on = os.environ.get("MMTG_GEN_TIMING") == "1"
        self.t = time.perf_counter()
        self.rows = []

    def mark(self, name):
        if self.on:
            torch.cuda.synchronize()
            now = time.perf_counter()
            self.rows.append(f"{name} {1e3 * (now - self.t):.2f} ms")
            self.t = now

    def done(self):
        if self.on:
            print("[mmtg generate] " + " | ".join(self.rows), flush=True)


@torch.no_grad()
def sample_sequence_batch(model, start_inputs, length, tokenizer=None, temperature=1.0, top_k=30, top_p=0.0,
                          repitition_penalty=1.0, device="cuda", seed=0, use_cuda_graph=True,
                          return_step_logits=False):
    """B independent runs of the reference's sample_sequence. `start_inputs`: list of per-sample
    dicts (MyDataset items with 'targets' = [start id]) or one dict of batched arrays.
    Returns a list of B token-id lists."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.MMTGError("mmtg_b200 generation runs on CUDA only (no CPU fallback)")
    lib = _lib.lib()
    tm = _SectionTimer()
    host = _collate(start_inputs)
    Bn = host["topic_ids"].shape[0]
    batch = {k: torch.as_tensor(host[k], dtype=torch.float32).to(dev) for k in _FLOAT_KEYS}
    for k in _INT_KEYS:
        batch[k] = torch.as_tensor(host[k]).long().to(dev)
    first = torch.as_tensor(host["targets"]).long().view(Bn, -1)[:, :1].to(dev)
    batch["targets"] = first
    batch["attention_mask"] = torch.ones_like(first)
    batch["type_ids"] = torch.zeros_like(first)
    dc = model.data_config
    sent_len = dc["max_sent_length"] + 2
    n_sent = dc["max_seq_length"] // sent_len
    tm.mark("collate+h2d")
    # ---- prefill: [prompt | first token] through the training-style engine (inference branch) ----
    prev_flag = model.train_flag
    model.train_flag = False
    try:
        _, _, logits0 = model(batch)
    finally:
        model.train_flag = prev_flag
    tm.mark("prefill")
    step = logits0._mmtg_step
    d = step.dims
    sampling = (float(temperature), int(top_k), float(top_p), float(repitition_penalty))
    # Persistent decode kernel for B <= 64: embedding + projector + all blocks + lm_head + sampler of
    # EVERY remaining position in ONE launch (mmtg_decode_steps_fused). Split-K partials are reduced
    # in a fixed order through cluster shared memory: bit-reproducible. MMTG_DECODE_MEGA=0 selects
    # the per-op launches (~90 kernels per position, the only path for B > 64).
    fused = (Bn <= 64 and d.E == 768 and d.P + length + 1 <= 1024 and d.He == 512 and 0 <= int(top_k) <= 1024
             and os.environ.get("MMTG_DECODE_MEGA", "1") != "0")
    key = (Bn, length, str(dev), sampling, model._flat[0].data_ptr(), model._table(dev).data_ptr(), fused)
    sessions = model.__dict__.setdefault("_decode_sessions", {})
    ses = sessions.get(key)
    if ses is None:
        if len(sessions) > 4:
            sessions.clear()
        ses = sessions[key] = _DecodeSession(model, d, length, Bn, dev, sampling)
    Lmax, gen_ld, dws, gen, j, step_logits = ses.Lmax, ses.gen_ld, ses.dws, ses.gen, ses.j, ses.step_logits
    st = C.c_void_p(_lib.stream_ptr())
    _lib.check(lib.mmtg_decode_load_prefix(C.byref(d), C.c_void_p(step.ws.data_ptr()), Lmax, C.c_void_p(dws.data_ptr()),
                                           C.c_void_p(step.mask.data_ptr()), st), "mmtg_decode_load_prefix")
    tm.mark("session+load_prefix")
    gen.zero_()
    gen[:, 0] = first[:, 0].to(torch.int32)
    j.zero_()
    ses.seed.fill_(int(seed))
    cm = model._c_model(d, dev)
    ses.cm = cm  # keep the struct alive for captured launches
    kept = []

    def sample(ptr, ld):
        _lib.check(lib.mmtg_sample_rows(C.c_void_p(ptr), C.c_int64(ld), C.c_void_p(gen.data_ptr()), gen_ld,
                                        C.c_void_p(j.data_ptr()), Bn, d.V, sent_len, C.c_float(temperature),
                                        int(top_k), C.c_float(top_p), C.c_float(repitition_penalty),
                                        C.c_uint64(0), C.c_void_p(ses.seed.data_ptr()), 1, None,
                                        C.c_void_p(_lib.stream_ptr())), "mmtg_sample_rows")

    if fused:  # LayerNorm-folded weight copies and the projector tables follow the current parameters / context
        _lib.check(lib.mmtg_decode_fold_weights(C.byref(cm), Lmax, C.c_void_p(dws.data_ptr()), st),
                   "mmtg_decode_fold_weights")

    def fused_steps(n):
        """n positions in one launch: consumes gen[:, j], decides gen[:, j+1 .. j+n], advances j by n."""
        _lib.check(lib.mmtg_decode_steps_fused(C.byref(cm), Lmax, C.c_void_p(dws.data_ptr()), C.c_void_p(gen.data_ptr()),
                                               gen_ld, C.c_void_p(j.data_ptr()), sent_len, n_sent, int(n),
                                               C.c_float(temperature), int(top_k), C.c_float(top_p),
                                               C.c_float(repitition_penalty), C.c_void_p(ses.seed.data_ptr()),
                                               C.c_void_p(step_logits.data_ptr()), C.c_void_p(_lib.stream_ptr())),
                   "mmtg_decode_steps_fused")

    # MMTG_DECODE_MEGA=2: the step-only entry of the persistent kernel (mmtg_decode_step_fused: blocks +
    # lm_head in the kernel; embedding / projector / sampler as separate launches) — kept for callers
    # that drive their own sampler, and cross-checked against the full-step mode in the tests
    step_only = fused and os.environ.get("MMTG_DECODE_MEGA", "1") == "2"
    step_fn = lib.mmtg_decode_step_fused if step_only else lib.mmtg_decode_step

    def one_step():
        if fused and not step_only:
            fused_steps(1)
            return
        _lib.check(step_fn(C.byref(cm), Lmax, C.c_void_p(dws.data_ptr()), C.c_void_p(gen.data_ptr()),
                                        gen_ld, C.c_void_p(j.data_ptr()), sent_len, n_sent,
                                        C.c_void_p(step_logits.data_ptr()), C.c_void_p(_lib.stream_ptr())),
                   "mmtg_decode_step")
        sample(step_logits.data_ptr(), d.V)

    tm.mark("fold")
    # reference iteration i = 0: logits of the last prefix row decide targets[1]
    if return_step_logits:
        kept.append(logits0[:, -1, :].clone())
    sample(logits0.data_ptr() + 4 * (d.L - 1) * d.V, d.L * d.V)
    remaining = length - 1
    ev = None
    if os.environ.get("MMTG_GEN_EVENTS") == "1":
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    graphed = use_cuda_graph and not return_step_logits and (not fused or step_only)
    if graphed and ses.graph is None and remaining > 2:
        for _ in range(2):  # eager warm-up (kernel attributes, descriptor cache), then capture
            l0 = _lib.launch_count()
            one_step()
            ses.launches_per_step = _lib.launch_count() - l0
        remaining -= 2
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                one_step()
        torch.cuda.current_stream().wait_stream(side)
        ses.graph = graph
    global replayed_launches, last_steps_ms
    if ev:
        ev[0].record()
    if fused and not step_only and not return_step_logits and remaining > 0:
        # use_cuda_graph=False: one launch per position (same kernel, n_steps = 1) instead of one per call
        l0 = _lib.launch_count()
        if use_cuda_graph:
            fused_steps(remaining)
        else:
            for _ in range(remaining):
                fused_steps(1)
        ses.launches_per_step = (_lib.launch_count() - l0) / remaining
    elif graphed and ses.graph is not None:
        for _ in range(remaining):
            ses.graph.replay()
        replayed_launches += remaining * getattr(ses, "launches_per_step", 0)
    else:
        for _ in range(remaining):
            one_step()
            if return_step_logits:
                kept.append(step_logits.clone())
    if ev:
        ev[1].record()
        ev[1].synchronize()
        last_steps_ms = ev[0].elapsed_time(ev[1]) * (length - 1) / max(1, remaining)
    tm.mark("steps")
    out = gen.cpu().numpy()
    tm.mark("d2h")
    tm.done()
    n_ret = returned_length(length, sent_len)
    rows = [out[b, :n_ret].tolist() for b in range(Bn)]
    if return_step_logits:
        return rows, kept
    return rows


def sample_sequence(model, start_input, length, tokenizer=None, temperature=1.0, top_k=30, top_p=0.0,
                    repitition_penalty=1.0, device="cuda", seed=0):
    """Reference signature (src/generate.py:97-107): one sample in, list of token ids out."""
    one = {k: np.asarray(v)[None] for k, v in start_input.items() if k != "rating"}
    return sample_sequence_batch(model, one, length, tokenizer, temperature, top_k, top_p, repitition_penalty,
                                 device, seed)[0]


def postprocess_tokens(tokens):
    """Detokenised generation -> lyric string (restates src/generate.py:222-236):
    cut after the 10th [#EOS#] (if no [SEP] precedes the last [#EOS#]) or at the first [SEP],
    drop [SEP] / [PAD] / [#START#], turn [#EOS#] into the Chinese comma and strip trailing commas.
    `tokens`: list of token strings (tokenizer.convert_ids_to_tokens of sample_sequence's ids)."""
    preds = list(tokens)
    eos = [i for i, v in enumerate(preds) if v == "[#EOS#]"]
    if len(eos) >= 10 and "[SEP]" not in preds[:eos[-1]]:
        preds = preds[:eos[9] + 1] + ["[SEP]"]
    elif "[SEP]" in preds:
        preds = preds[:preds.index("[SEP]") + 1]
    else:
        preds = preds + ["[SEP]"]
    text = "".join(preds).replace("[SEP]", "").replace("[PAD]", "").replace("[#START#]", "").replace("[#EOS#]", "，")
    while text and text[-1] == "，":
        text = text[:-1]
    return text


def generate_samples(model, dataset, tokenizer, n_samples=10, length=None, start_token="[#START#]", temperature=1.1,
                     top_k=10, top_p=0.7, repitition_penalty=1.5, device="cuda", rows_per_call=64, seed=0,
                     save_path=None, _sampler=None):
    """The reference's generation front-end (src/generate.py:203-244) over the batched decoder:
    for every dataset item, `n_samples` independent runs of `sample_sequence` started from
    `[#START#]`, detokenised and post-processed with `postprocess_tokens`; one output line per
    sample, items in dataset order (written to `save_path` when given). The (item, sample) runs
    are independent batch-1 reference runs, so they are packed `rows_per_call` at a time into
    `sample_sequence_batch` (the reference decodes them one after another).
    `dataset[i]` is a MyDataset item (dict of arrays); `tokenizer` needs convert_tokens_to_ids /
    convert_ids_to_tokens. Returns the list of strings."""
    if length is None:
        length = model.data_config["max_seq_length"]
    start_id = tokenizer.convert_tokens_to_ids(start_token)
    sampler = sample_sequence_batch if _sampler is None else _sampler
    jobs = [(i, s) for i in range(len(dataset)) for s in range(n_samples)]
    out = [None] * len(jobs)
    for lo in range(0, len(jobs), rows_per_call):
        chunk = jobs[lo:lo + rows_per_call]
        starts = []
        for i, _ in chunk:
            item = {k: np.asarray(v) for k, v in dataset[i].items() if k != "rating"}
            item["targets"] = np.asarray([start_id])
            starts.append(item)
        rows = sampler(model, starts, length, tokenizer, temperature, top_k, top_p, repitition_penalty, device,
                       seed + lo)
        for k, ids in enumerate(rows):
            out[lo + k] = postprocess_tokens(tokenizer.convert_ids_to_tokens(ids))
    if save_path is not None:
        with open(save_path, "w", encoding="utf-8") as f:
            for line in out:
                f.write(line + "\n")
    return out
