"""GPU parity at the BENCHMARKED configuration (BASELINE.json config #2: C = 256, 8 heads of 32 channels, 100 queries,
9 decoder layers, key grids 15x20 / 30x40 / 60x80, masks 120x160) - the shapes at which the decoder takes the tcgen05
attention kernel (the small golden fixtures have 16-channel heads and run the CUDA-core kernel).

What "parity" can mean for this model: every decoder layer thresholds the previous layer's masks
(``sigmoid < 0.5``), so two correct fp32 implementations that differ by 1e-6 occasionally disagree on a mask bit, after
which the layers behind it see different inputs (the reference's own CPU-vs-CPU runs do this too). Hence:

* teacher-forced (each layer gets the ORACLE's input state and mask bits): every layer's logits must agree to <= 1e-3 of
  peak (north_star's tolerance; measured ~1e-5), mask bits and per-pixel labels must be bit-exact outside a stated margin;
* free-running over 8 seeds: the statistics are asserted and recorded (layers before the first flipped bit agree to
  1e-4; label agreement per seed; median of the final error);
* the TIMED configuration (B = 8, one CUDA graph, programmatic dependent launch on) must reproduce the eager forward
  bit for bit;
* bounded random shape sweeps of the dense / attention kernels against fp64.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import decoder as odec

pytestmark = pytest.mark.gpu

LEVELS = [(15, 20), (30, 40), (60, 80)]
MASK_HW = (120, 160)
HEADS, LAYERS, Q, C = 8, 9, 100, 256


def peak_rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _pack_bits(blocked):
    B, Qn, S = blocked.shape
    words = (S + 31) // 32
    pad = torch.zeros(B, Qn, words * 32, dtype=torch.bool, device=blocked.device)
    pad[..., :S] = blocked
    v = (pad.view(B, Qn, words, 32).long() << torch.arange(32, device=blocked.device)).sum(-1)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


def _decoder(seed):
    from unseenobjectswithmeanshift_b200 import workloads
    from unseenobjectswithmeanshift_b200.meanshiftformer import modeling as M
    torch.manual_seed(seed)
    m = M.MeanShiftTransformerDecoder(64, True, **workloads.decoder_kwargs(LAYERS)).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m, sd


def _inputs(batch, seed):
    g = torch.Generator().manual_seed(7000 + seed)
    x = [torch.randn(batch, 64, h, w, generator=g) for h, w in LEVELS]
    mf = torch.randn(batch, C, *MASK_HW, generator=g)
    return x, mf


@pytest.mark.parametrize("block_kernel", ["0", "1"])
def test_decoder_config2_teacher_forced(monkeypatch, block_kernel):
    """Each layer of the CUDA decoder is fed the oracle's own layer input (query state + mask bits); its prediction
    (class logits, mask logits) and the NEXT layer's mask bits it derives are compared with the oracle's. Both forms of
    the layer: the per-op kernels (default) and the one-launch cluster kernel (MSM_DECODER_BLOCK=1)."""
    from unseenobjectswithmeanshift_b200 import ops
    monkeypatch.setenv("MSM_DECODER_BLOCK", block_kernel)
    B = 2
    m, sd = _decoder(0)
    x, mf = _inputs(B, 0)
    trace = []
    with torch.no_grad():
        ref = odec.decoder_forward(sd, x, mf, num_heads=HEADS, num_layers=LAYERS, trace=trace)
    ref_masks = [a["pred_masks"] for a in ref["aux_outputs"]] + [ref["pred_masks"]]
    ref_logits = [a["pred_logits"] for a in ref["aux_outputs"]] + [ref["pred_logits"]]
    m = m.cuda()
    bit_mismatch = []

    def teacher(i, out, bits, row_open):
        S = LEVELS[i % 3][0] * LEVELS[i % 3][1]
        want = trace[i]["blocked"].view(B, HEADS, Q, S)[:, 0]          # identical for the 8 heads, after the :618 rule
        got = ops.unpack_attn_bits(bits, row_open, S, 1).view(B, Q, S).cpu()
        bit_mismatch.append((got != want).float().mean().item())
        forced_bits = _pack_bits(want.cuda())
        return (trace[i]["tgt_in"].transpose(0, 1).contiguous().cuda(), forced_bits,
                torch.ones(B, Q, dtype=torch.int32, device="cuda"))

    with torch.no_grad():
        out = m([t.cuda() for t in x], mf.cuda(), _teacher=teacher)
    got_masks = [a["pred_masks"].cpu() for a in out["aux_outputs"]] + [out["pred_masks"].cpu()]
    got_logits = [a["pred_logits"].cpu() for a in out["aux_outputs"]] + [out["pred_logits"].cpu()]
    errs = [peak_rel(g, r) for g, r in zip(got_masks, ref_masks)]
    lerrs = [peak_rel(g, r) for g, r in zip(got_logits, ref_logits)]
    print("teacher-forced mask-logit error per prediction (of peak):", ["%.1e" % e for e in errs])
    print("teacher-forced class-logit error per prediction (of peak):", ["%.1e" % e for e in lerrs])
    print("mask-bit mismatch fraction per layer:", ["%.1e" % e for e in bit_mismatch])
    assert max(errs) < 1e-3 and max(lerrs) < 1e-3              # north_star: 1e-3 rel fp32 on mask logits
    assert max(errs) < 1e-4, "split-precision tensor-core path is expected at ~1e-5"
    # mask bits: a bit may differ only where the oracle's own resampled logit is within 1e-4 of peak of the threshold;
    # at 19200..4800 keys x 100 queries x 2 images that is a handful of bits at most
    assert max(bit_mismatch) < 2e-5
    # per-pixel instance labels (argmax over queries), bit-exact wherever the oracle's top-2 margin exceeds the tolerance
    for g, r in zip(got_masks, ref_masks):
        top2 = r.topk(2, dim=1).values
        decided = (top2[:, 0] - top2[:, 1]) > 1e-4 * r.abs().max()
        same = g.argmax(1) == r.argmax(1)
        assert bool(same[decided].all())
        assert same.float().mean().item() > 0.9995


def test_decoder_config2_free_running_statistics():
    """No teacher: 8 seeds (weights and inputs), one image each. Asserts what holds for every seed and records the
    statistics the design doc quotes (gpurun_out/parity_config2_seeds.json when the directory exists)."""
    rows = []
    for seed in range(8):
        m, sd = _decoder(seed)
        x, mf = _inputs(1, seed)
        with torch.no_grad():
            ref = odec.decoder_forward(sd, x, mf, num_heads=HEADS, num_layers=LAYERS)
            out = m.cuda()([t.cuda() for t in x], mf.cuda())
        ref_masks = [a["pred_masks"] for a in ref["aux_outputs"]] + [ref["pred_masks"]]
        got_masks = [a["pred_masks"].cpu() for a in out["aux_outputs"]] + [out["pred_masks"].cpu()]
        errs = [peak_rel(g, r) for g, r in zip(got_masks, ref_masks)]
        first = next((i for i, e in enumerate(errs) if e > 1e-4), None)   # prediction index of the first visible flip
        r, g = ref_masks[-1], got_masks[-1]
        rows.append({"seed": seed, "first_flip_prediction": first, "final_err_of_peak": errs[-1],
                     "frac_gt_1e-3": ((g - r).abs() > 1e-3 * r.abs().max()).float().mean().item(),
                     "argmax_agreement": (g.argmax(1) == r.argmax(1)).float().mean().item(),
                     "per_prediction_err": errs})
        # before any bit flips the two paths agree at kernel accuracy
        upto = len(errs) if first is None else first
        assert all(e < 1e-4 for e in errs[:upto])
        assert errs[0] < 2e-5 and errs[1] < 1e-4
    print(json.dumps(rows))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_config2_seeds.json"), "w") as f:
            json.dump(rows, f, indent=1)
    finals = sorted(r["final_err_of_peak"] for r in rows)
    assert finals[len(finals) // 2] < 1e-3, finals                      # median over seeds inside north_star's 1e-3
    assert min(r["argmax_agreement"] for r in rows) > 0.97, rows         # a flipped bit moves a few % of labels at most
    assert sum(r["argmax_agreement"] > 0.9995 for r in rows) >= 4, rows  # most seeds: no visible flip at all


def test_graph_replay_equals_eager_config2():
    """The timed configuration of bench.py: B = 8, the whole R50-config head as ONE CUDA graph with programmatic
    dependent launch on (the default). Replay must reproduce the eager forward bit for bit, twice in a row (static
    buffers reused), and a fresh input must give fresh results."""
    from unseenobjectswithmeanshift_b200 import workloads
    from unseenobjectswithmeanshift_b200.graph import GraphedForward
    assert os.environ.get("MSM_DISABLE_PDL", "") in ("", "0")
    head = workloads.build_head("r50", seed=0).cuda()
    feats = {k: v.cuda() for k, v in workloads.synthetic_features("r50", 8, seed=0).items()}
    feats2 = {k: v.cuda() for k, v in workloads.synthetic_features("r50", 8, seed=1).items()}

    def fwd(f):
        out, _ = head(f, 480, 640)
        return {"pred_logits": out["pred_logits"], "pred_masks": out["pred_masks"],
                "aux": [a["pred_masks"] for a in out["aux_outputs"]]}

    with torch.no_grad():
        eager = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v]) for k, v in fwd(feats).items()}
        eager2 = fwd(feats2)["pred_masks"].clone()
    g = GraphedForward(fwd, feats)
    for _ in range(2):
        out = g(feats)
        torch.cuda.synchronize()
        assert torch.equal(out["pred_masks"], eager["pred_masks"])
        assert torch.equal(out["pred_logits"], eager["pred_logits"])
        for a, b in zip(out["aux"], eager["aux"]):
            assert torch.equal(a, b)
    out = g(feats2)
    torch.cuda.synchronize()
    assert torch.equal(out["pred_masks"], eager2)
    assert not torch.equal(eager2, eager["pred_masks"])


# ----------------------------------------------------------------------------- bounded random shape sweeps
def _rng(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("seed", range(24))
def test_linear_shape_sweep_vs_fp64(seed):
    """msm_linear_fwd on random (M, N, K): N, K multiples of 32, ragged M (row tiles with tails, single rows, several
    waves), optional bias / ReLU - every element within 2e-5 of the output's peak of the fp64 result."""
    from unseenobjectswithmeanshift_b200 import ops
    g = _rng(100 + seed)
    cpu = torch.Generator().manual_seed(100 + seed)
    M = int(torch.randint(1, 40000 if seed % 4 == 0 else 3000, (1,), generator=cpu))
    N = 32 * int(torch.randint(1, 33 if seed % 3 else 9, (1,), generator=cpu))
    K = 32 * int(torch.randint(1, 65 if seed % 5 == 0 else 17, (1,), generator=cpu))
    relu, has_bias = bool(seed & 1), bool(seed & 2)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g) if has_bias else None
    y = ops.linear(x, w, b, relu=relu)
    ref = x.double() @ w.double().t() + (b.double() if has_bias else 0.0)
    ref = ref.relu() if relu else ref
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, (M, N, K, relu, has_bias, err)


@pytest.mark.parametrize("seed", range(16))
def test_linear_ln_and_ffn_shape_sweep_vs_fp64(seed):
    """msm_linear_ln_fwd (N in {32, 64}: the row epilogue must see whole rows whatever the planner does with few row
    tiles - the bug class of round 1) and msm_ffn_ln_fwd on random row counts / widths."""
    from unseenobjectswithmeanshift_b200 import ops
    g = _rng(200 + seed)
    cpu = torch.Generator().manual_seed(200 + seed)
    M = int(torch.randint(1, 20000 if seed % 4 == 0 else 1500, (1,), generator=cpu))
    N = 32 * int(torch.randint(1, 3, (1,), generator=cpu))
    K = 32 * int(torch.randint(1, 40, (1,), generator=cpu))
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    norm = torch.nn.LayerNorm(N).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(N, device="cuda", generator=g))
        norm.bias.copy_(torch.randn(N, device="cuda", generator=g))
        y = ops.linear_ln(x, w, b, res, norm)
        ref = F.layer_norm(res.double() + x.double() @ w.double().t() + b.double(), (N,), norm.weight.double(),
                           norm.bias.double(), norm.eps)
    assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < 3e-5, (M, N, K)
    D = N
    Fd = 128 * int(torch.randint(1, 15, (1,), generator=cpu))
    xx = torch.randn(M, D, device="cuda", generator=g)
    w1 = torch.randn(Fd, D, device="cuda", generator=g) / D ** 0.5
    b1 = torch.randn(Fd, device="cuda", generator=g)
    w2 = torch.randn(D, Fd, device="cuda", generator=g) / Fd ** 0.5
    b2 = torch.randn(D, device="cuda", generator=g)
    with torch.no_grad():
        y = ops.ffn_ln(xx, w1, b1, w2, b2, norm)
        h = (xx.double() @ w1.double().t() + b1.double()).relu()
        ref = F.layer_norm(xx.double() + h @ w2.double().t() + b2.double(), (D,), norm.weight.double(),
                           norm.bias.double(), norm.eps)
    assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < 3e-5, (M, D, Fd)


@pytest.mark.parametrize("seed", range(24))
def test_attention_shape_sweep_vs_fp64(seed):
    """vMF attention on random (B, H, Q <= 128, S, hd in {32, 64}), with and without bit masks (some rows fully
    blocked -> the un-mask rule), through BOTH tcgen05 kernels: fp32 K / V rows (msm_vmf_attention_fwd) and packed
    operand images (TMA-streamed, the decoder's default), against the fp64 formula of attention_util.py:64-82."""
    from unseenobjectswithmeanshift_b200 import ops
    g = _rng(300 + seed)
    cpu = torch.Generator().manual_seed(300 + seed)
    B = int(torch.randint(1, 4, (1,), generator=cpu))
    H = int(torch.randint(1, 9, (1,), generator=cpu))
    Qn = int(torch.randint(1, 129, (1,), generator=cpu))
    S = int(torch.randint(1, 20000 if seed % 6 == 0 else 1500, (1,), generator=cpu))
    hd = 32 if seed % 3 else 64
    masked = bool(seed & 1)
    Cc = H * hd
    q = torch.randn(B, Qn, Cc, device="cuda", generator=g)
    kv = torch.randn(B, S, 2 * Cc, device="cuda", generator=g)
    hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)  # noqa: E731
    bits = ro = None
    blocked = torch.zeros(B, Qn, S, dtype=torch.bool, device="cuda")
    if masked:
        blocked = torch.rand(B, Qn, S, device="cuda", generator=g) < 0.6
        blocked[:, Qn // 2] = True
        ro = (~blocked).any(-1).to(torch.int32).contiguous()
        bits = _pack_bits(blocked)
        blocked = blocked & (ro != 0).unsqueeze(-1)
    q4, k4, v4 = hv(q), hv(kv[..., :Cc]), hv(kv[..., Cc:])
    qn, kn = F.normalize(q4.double(), dim=-1), F.normalize(k4.double(), dim=-1)
    s = 30.0 * qn @ kn.transpose(-1, -2)
    s = s.masked_fill(blocked.unsqueeze(1), float("-inf"))
    ref = F.normalize(torch.softmax(s, -1) @ v4.double(), dim=-1)
    got = ops.vmf_attention(q4, k4, v4, blocked_bits=bits, row_open=ro)
    assert (got.double() - ref).abs().max().item() < 1e-4, ("rows", B, H, Qn, S, hd, masked)
    if hd == 32:
        got = ops.vmf_attention_packed(q4, ops.pack_kv(k4, v4), blocked_bits=bits, row_open=ro)
        assert (got.double() - ref).abs().max().item() < 1e-4, ("packed", B, H, Qn, S, hd, masked)


# ----------------------------------------------------------------------------- a6, narrow heads
@pytest.mark.parametrize("hw", [(15, 20), (60, 80), (224, 224), (7, 5)])
def test_position_embedding_sine_table_vs_oracle(hw):
    """SURVEY 8 row a6 directly: the cached separable [S, C] table (and the reference-shaped forward) against the
    oracle's cumsum formulation of position_encoding.py:29-52, normalize=True, scale 2 pi."""
    from unseenobjectswithmeanshift_b200.meanshiftformer.modeling.transformer_decoder.position_encoding import \
        PositionEmbeddingSine
    h, w = hw
    pe = PositionEmbeddingSine(128, normalize=True)
    want = odec.position_embedding_sine(2, h, w, 128)                       # [B, 256, h, w]
    got = pe(torch.zeros(2, 3, h, w, device="cuda")).cpu()
    tab = pe.table(h, w, torch.device("cuda")).cpu()                        # [h*w, 256]
    assert got.shape == want.shape
    # sin / cos of arguments up to 2 pi: the device's and the host's libm agree to a few ulp
    assert (got - want).abs().max().item() < 2e-6
    assert (tab - want[0].flatten(1).t()).abs().max().item() < 2e-6


def test_narrow_head_runs_in_the_library_gemm():
    """The 3-way class head (N = K + 1 = 3): zero-padded to 32 columns, same tensor-core kernel, no cuBLAS."""
    from unseenobjectswithmeanshift_b200 import ops
    g = _rng(9)
    x = torch.randn(8, 100, 256, device="cuda", generator=g)
    w = torch.randn(3, 256, device="cuda", generator=g) / 16
    b = torch.randn(3, device="cuda", generator=g)
    ops.reset_stats()
    y = ops.dense(x, w, b)
    assert ops.launches() == 1 and y.shape == (8, 100, 3)
    ref = x.double() @ w.double().t() + b.double()
    assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < 2e-5
    with torch.no_grad():
        w.mul_(2.0)   # in-place edit: the padded copy must follow
    y2 = ops.dense(x, w, b)
    ref2 = x.double() @ w.double().t() + b.double()
    assert (y2.double() - ref2).abs().max().item() / ref2.abs().max().item() < 2e-5


# ----------------------------------------------------------------------------- decoder-layer cluster kernel
@pytest.mark.parametrize("B,Qn,last,block_norm", [(8, 100, False, True), (1, 100, True, True), (3, 128, False, False),
                                                   (2, 37, False, True)])
def test_decoder_block_kernel_vs_fp64(B, Qn, last, block_norm):
    """msm_decoder_block_fwd (cluster of 8 CTAs per image: out_proj + LN, self-attention block, FFN block, block norm,
    decoder_norm, class head, mask MLP, next layer's query projection) against the same chain in fp64 torch ops,
    following meanshiftformer_transformer_decoder.py:171-181, 253-257, 300-304, 637-638, 661-664."""
    from unseenobjectswithmeanshift_b200 import ops
    C, H, FF = 256, 8, 2048
    g = _rng(900 + B + Qn)
    rn = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc  # noqa: E731
    W = dict(o1=rn(C, C, sc=C ** -0.5), qkv=rn(3 * C, C, sc=C ** -0.5), o2=rn(C, C, sc=C ** -0.5),
             f1=rn(FF, C, sc=C ** -0.5), f2=rn(C, FF, sc=FF ** -0.5), qn=rn(C, C, sc=C ** -0.5),
             m1=rn(C, C, sc=C ** -0.5), c=rn(3, C, sc=C ** -0.5), m2=rn(C, C, sc=C ** -0.5), m3=rn(C, C, sc=C ** -0.5))
    bvec = {k: rn(v.shape[0], sc=0.1) for k, v in W.items()}
    norms = []
    for _ in range(4):
        n = torch.nn.LayerNorm(C).cuda()
        with torch.no_grad():
            n.weight.copy_(1 + rn(C, sc=0.2))
            n.bias.copy_(rn(C, sc=0.2))
        norms.append(n)
    qpos = rn(Qn, C)
    o = F.normalize(rn(B, Qn, H, C // H), dim=-1).reshape(B, Qn, C)    # a cross-attention output: unit rows per head
    state = rn(B, Qn, C)
    with torch.no_grad():
        blob = ops.decoder_block_pack(W["o1"], W["qkv"], W["o2"], W["f1"], W["f2"], None if last else W["qn"], W["m1"],
                                      W["c"], W["m2"], W["m3"])
        tqk = torch.cat([qpos @ W["qkv"][:2 * C].t(), qpos.new_zeros(Qn, C)], 1).contiguous()
        tqn = (qpos @ W["qn"].t()).contiguous()
        bc32 = torch.cat([bvec["c"], bvec["c"].new_zeros(29)])
        got = ops.decoder_block(o, state, blob, b_o1=bvec["o1"], norm1=norms[0], b_qkv=bvec["qkv"], t_qk=tqk,
                                b_o2=bvec["o2"], norm2=norms[1], b_f1=bvec["f1"], b_f2=bvec["f2"], norm3=norms[2],
                                block_norm=block_norm, normd=norms[3], b_qn=None if last else bvec["qn"],
                                t_qn=None if last else tqn, b_m1=bvec["m1"], b_c32=bc32, b_m2=bvec["m2"],
                                b_m3=bvec["m3"])
        torch.cuda.synchronize()
        d = lambda t: t.double()  # noqa: E731
        ln = lambda x, n: F.layer_norm(x, (C,), d(n.weight), d(n.bias), n.eps)  # noqa: E731
        t1 = ln(d(state) + d(o) @ d(W["o1"]).t() + d(bvec["o1"]), norms[0])
        qk = t1 + d(qpos)
        q = qk @ d(W["qkv"][:C]).t() + d(bvec["qkv"][:C])
        k = qk @ d(W["qkv"][C:2 * C]).t() + d(bvec["qkv"][C:2 * C])
        v = t1 @ d(W["qkv"][2 * C:]).t() + d(bvec["qkv"][2 * C:])
        hv = lambda t: t.unflatten(-1, (H, C // H)).transpose(1, 2)  # noqa: E731
        a = torch.softmax(30.0 * F.normalize(hv(q), dim=-1) @ F.normalize(hv(k), dim=-1).transpose(-1, -2), -1) @ hv(v)
        a = F.normalize(a, dim=-1).transpose(1, 2).reshape(B, Qn, C)
        t2 = ln(t1 + a @ d(W["o2"]).t() + d(bvec["o2"]), norms[1])
        t3 = ln(t2 + (t2 @ d(W["f1"]).t() + d(bvec["f1"])).relu() @ d(W["f2"]).t() + d(bvec["f2"]), norms[2])
        if block_norm:
            t3 = F.normalize(t3, dim=-1)
        dec = ln(t3, norms[3])
        logits = dec @ d(W["c"]).t() + d(bvec["c"])
        e = (dec @ d(W["m1"]).t() + d(bvec["m1"])).relu()
        e = (e @ d(W["m2"]).t() + d(bvec["m2"])).relu()
        embed = e @ d(W["m3"]).t() + d(bvec["m3"])
        qn = (t3 + d(qpos)) @ d(W["qn"]).t() + d(bvec["qn"])
    state_out, logits32, emb, q_next = got
    for name, x, r in (("state", state_out, t3), ("logits", logits32[..., :3], logits), ("embed", emb, embed)):
        err = (d(x) - r).abs().max().item() / r.abs().max().item()
        assert err < 3e-5, (name, err)
    if last:
        assert q_next is None
    else:
        assert (d(q_next) - qn).abs().max().item() / qn.abs().max().item() < 3e-5


def test_lean_eval_path_matches_full_masks():
    """eval_aux_masks=False (what the META_ARCH wrappers set: the eval branch never reads aux_outputs): intermediate
    layers compute their mask logits on the next layer's key grid only - interpolate(einsum(e, F)) == einsum(e,
    interpolate(F)). The final prediction must agree with the full path; an intermediate prediction's low-resolution
    logits must equal the bilinear resample of the full path's logits."""
    m, _ = _decoder(3)
    m = m.cuda()
    x, mf = _inputs(2, 3)
    xc, mfc = [t.cuda() for t in x], mf.cuda()
    with torch.no_grad():
        full = m(xc, mfc)
        m.eval_aux_masks = False
        lean = m(xc, mfc)
        m.eval_aux_masks = True
    assert lean["pred_masks"].shape == full["pred_masks"].shape
    # prediction 0 (from the learnable queries) feeds layer 0 at the 15x20 grid
    lo = lean["aux_outputs"][0]["pred_masks"]
    assert tuple(lo.shape[-2:]) == LEVELS[0]
    want = F.interpolate(full["aux_outputs"][0]["pred_masks"], size=LEVELS[0], mode="bilinear", align_corners=False)
    assert peak_rel(lo, want) < 2e-5
    # the two paths may disagree on a mask bit whose logit is within 1e-6 of zero; what they must not do is differ visibly
    assert peak_rel(lean["pred_logits"], full["pred_logits"]) < 1e-2
    agree = (lean["pred_masks"].argmax(1) == full["pred_masks"].argmax(1)).float().mean().item()
    assert agree > 0.995, agree
    first = peak_rel(lean["aux_outputs"][1]["pred_logits"], full["aux_outputs"][1]["pred_logits"])
    assert first < 1e-4, first   # after one layer (before any flip can matter much) the class logits are identical


@pytest.mark.parametrize("shape,size", [((2, 256, 120, 160), (15, 20)), ((2, 256, 120, 160), (60, 80)),
                                        ((1, 3, 7, 5), (13, 9)), ((3, 5, 33, 17), (33, 17)), ((1, 2, 1, 1), (4, 6)),
                                        ((2, 100, 120, 160), (480, 640))])
def test_resample_bilinear_vs_interpolate(shape, size):
    """msm_resample_bilinear_fwd == F.interpolate(bilinear), both align_corners rules: down-, up-sampling, identity, 1x1."""
    from unseenobjectswithmeanshift_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, generator=g).cuda()
    for ac in (False, True):
        got = ops.resample_bilinear(x, size, align_corners=ac)
        want = F.interpolate(x, size=size, mode="bilinear", align_corners=ac)
        assert got.shape == want.shape
        assert (got - want).abs().max().item() <= 4e-6 * max(1.0, want.abs().max().item()), ac


def test_segnet_embedding_single_upsample_matches_two():
    """backbones.SegnetEmbedding on the GPU sums the RGB and depth streams at 1/8 resolution and up-samples once
    (bilinear, align_corners=True, is linear) - against the module's plain path (two up-samplings, then the sum)."""
    from unseenobjectswithmeanshift_b200 import backbones
    backbones.set_tf32(False)
    try:
        m = backbones.SegnetEmbedding(seed=1).cuda()
        g = torch.Generator().manual_seed(3)
        img, dep = torch.randn(1, 3, 96, 128, generator=g).cuda(), torch.randn(1, 3, 96, 128, generator=g).cuda()
        with torch.no_grad(), backbones._conv_math(False):   # (the module sets its conv math itself; the streams do not)
            got = m(img, None, dep)
            want = m.fcn(img) + m.fcn_depth(dep)
        assert got.is_contiguous() and got.shape == want.shape
        assert (got - want).abs().max().item() / want.abs().max().item() < 1e-5
    finally:
        backbones.set_tf32(True)


@pytest.mark.parametrize("B,Hh,Ww,layers,masked", [(2, 16, 64, 2, True),    # image rows never wrap inside a warp
                                                    (1, 24, 40, 3, True),    # they do: per-lane ty loads
                                                    (2, 30, 20, 1, False),   # S = 600: last key tile has a tail
                                                    (1, 5, 7, 2, True)])     # S % 4 != 0: token-major x, tables only
def test_folded_projection_to_packed_attention_vs_fp64(B, Hh, Ww, layers, masked):
    """The K / V projections of the UCN / crop configs with input_proj folded in (K = 64): x [B, 64, H, W] channel-major,
    K = x (W_k W_in)^T + b + ty[y] + tx[x] (separable sine tables in the epilogue), V = x (W_v W_in)^T + b, written as
    operand images, then the packed attention per layer - against the fp64 cross-attention on
    K = (W_in x + b_in + pos) W_k^T + b_k, V = (W_in x + b_in) W_v^T + b_v (attention_util.py:64-82, :121-140)."""
    from unseenobjectswithmeanshift_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(B * 1000 + Hh * Ww + layers)
    Cin, C, Hd, Q, S = 64, 256, 8, 100, Hh * Ww
    x = torch.randn(B, Cin, Hh, Ww, generator=g)
    w_in, b_in = torch.randn(C, Cin, generator=g) / Cin ** 0.5, torch.randn(C, generator=g) * 0.1
    wk, bk = torch.randn(layers * C, C, generator=g) / C ** 0.5, torch.randn(layers * C, generator=g) * 0.1
    wv, bv = torch.randn(layers * C, C, generator=g) / C ** 0.5, torch.randn(layers * C, generator=g) * 0.1
    ty, tx = torch.randn(Hh, C // 2, generator=g), torch.randn(Ww, C // 2, generator=g)   # any separable embedding
    q = torch.randn(layers, B, Q, C, generator=g)
    d = lambda t: t.double()  # noqa: E731
    # fp64 reference, the long way
    pos = torch.cat((ty[:, None, :].expand(Hh, Ww, -1), tx[None, :, :].expand(Hh, Ww, -1)), dim=2).reshape(S, C)
    src = d(x).flatten(2).transpose(1, 2) @ d(w_in).t() + d(b_in)
    Kref = ((src + d(pos)) @ d(wk).t() + d(bk)).view(B, S, layers, Hd, 32)
    Vref = (src @ d(wv).t() + d(bv)).view(B, S, layers, Hd, 32)
    # folded weights (what the decoder caches)
    fold_w = lambda w_: (d(w_) @ d(w_in)).float().contiguous().to(dev)  # noqa: E731
    fold_b = lambda w_, b_: (d(w_) @ d(b_in) + d(b_)).float().contiguous().to(dev)  # noqa: E731
    npf = C // 2
    tabs = ((d(ty) @ d(wk)[:, :npf].t()).float().contiguous().to(dev), (d(tx) @ d(wk)[:, npf:].t()).float().contiguous().to(dev))
    images, per_layer = ops.packed_kv_alloc(layers, B, Hd, S, dev)
    xin = x.to(dev).contiguous() if S % 4 == 0 else x.to(dev).flatten(2).transpose(1, 2).contiguous()
    ops.linear_packed_kv(xin, fold_w(wk), fold_b(wk, bk), images, B, S, C, 0, pos=tabs)
    ops.linear_packed_kv(xin, fold_w(wv), fold_b(wv, bv), images, B, S, C, 1)
    for j in range(layers):
        bits = ro = eff = None
        if masked:
            blocked = torch.rand(B, Q, S, generator=g) < 0.5
            blocked[:, 3] = True
            ro = (~blocked).any(-1).to(torch.int32).contiguous().to(dev)
            words = (S + 31) // 32
            pad = torch.zeros(B, Q, words * 32, dtype=torch.bool)
            pad[..., :S] = blocked
            v = (pad.view(B, Q, words, 32).long() << torch.arange(32)).sum(-1)
            bits = torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous().to(dev)
            eff = (blocked & (ro.cpu() != 0).unsqueeze(-1)).unsqueeze(1)
        q4 = q[j].to(dev).unflatten(-1, (Hd, 32)).permute(0, 2, 1, 3)
        kv = ops.PackedKV(images[j * per_layer:(j + 1) * per_layer], B, Hd, S)
        out = ops.vmf_attention_packed(q4, kv, blocked_bits=bits, row_open=ro)
        kj, vj = Kref[:, :, j].permute(0, 2, 1, 3), Vref[:, :, j].permute(0, 2, 1, 3)
        s = 30.0 * F.normalize(d(q[j]).unflatten(-1, (Hd, 32)).permute(0, 2, 1, 3), dim=-1) @ F.normalize(kj, dim=-1).transpose(-1, -2)
        if eff is not None:
            s = s.masked_fill(eff, float("-inf"))
        ref = F.normalize(torch.softmax(s, -1) @ vj, dim=-1)
        err = (out.cpu().double() - ref).abs().max().item()
        assert err == err and err < 3e-5, (j, err)


@pytest.mark.parametrize("B,K,N,Hh,Ww,bias", [(2, 256, 256, 16, 24, True),    # H*W = 384 = 3 x 128: channel-major result
                                              (1, 512, 256, 15, 20, True),    # 300 pixels: token-major result, NCHW view
                                              (2, 64, 64, 8, 16, False)])
def test_conv1x1_channels_last_input_vs_fp64(B, K, N, Hh, Ww, bias):
    """A channels_last feature map (what the cuDNN channels_last backbone hands over) through the 1x1 convolutions of
    the pixel decoder without a layout copy: msm_conv1x1_nhwc_fwd / the token-major linear, against fp64 conv2d."""
    from unseenobjectswithmeanshift_b200 import ops
    g = torch.Generator().manual_seed(B + K + Hh)
    x = torch.randn(B, K, Hh, Ww, generator=g)
    conv = torch.nn.Conv2d(K, N, 1, bias=bias)
    with torch.no_grad():
        want = F.conv2d(x.double(), conv.weight.double(), conv.bias.double() if bias else None)
        xc = x.cuda().contiguous(memory_format=torch.channels_last)
        assert ops.conv1x1_nhwc_supported(xc, conv.weight.cuda())
        got = ops.conv1x1_layer(conv.cuda(), xc)
    assert got.shape == want.shape
    assert (got.cpu().double() - want).abs().max().item() / want.abs().max().item() < 1e-5   # split-precision products


def test_backbone_fused_channels_last_matches_plain_module():
    """backbones.ResNet50Features: BatchNorms folded + cuDNN fused conv/bias/ReLU + downsample bias folded into conv3 +
    channels_last outputs == the plain torchvision module with eval-mode BatchNorms (randomised statistics), at fp32
    conv math."""
    import copy
    from unseenobjectswithmeanshift_b200 import backbones
    backbones.set_tf32(True)     # the timed configuration's layout (channels_last) ...
    plain = backbones.ResNet50Features(seed=3, fold_bn=False)
    plain.tf32 = False           # ... at fp32 math, so that the comparison is tight
    g = torch.Generator().manual_seed(1)
    for m in plain.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    fused = copy.deepcopy(plain).fold_()
    plain, fused = plain.cuda(), fused.cuda()
    x = torch.randn(2, 3, 64, 96, generator=g).cuda()
    with torch.no_grad():
        want, got = plain(x), fused(x)
    assert fused.fused_relu and not plain.fused_relu
    for k in want:
        assert got[k].shape == want[k].shape
        assert got[k].permute(0, 2, 3, 1).is_contiguous()      # handed to the head without an NCHW copy
        scale = want[k].abs().max().item()
        assert (got[k] - want[k]).abs().max().item() / scale < 1e-4, k


def test_maxpool_and_upsample_add_vs_torch():
    """The two layout-aware helpers around the pixel decoder: msm_maxpool3x3s2_nhwc_fwd == nn.MaxPool2d(3, 2, 1) on a
    channels_last map (bit-exact), msm_upsample_add_fwd == cur + F.interpolate(x, size, bilinear)."""
    from unseenobjectswithmeanshift_b200 import ops
    g = torch.Generator().manual_seed(2)
    for shape in ((2, 64, 30, 44), (1, 8, 7, 9), (2, 4, 1, 5)):
        x = torch.randn(*shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)
        got = ops.maxpool3x3s2_channels_last(x)
        want = F.max_pool2d(x, 3, 2, 1)
        assert got.shape == want.shape and torch.equal(got, want)
    for (B, C, h, w), (Ht, Wt) in (((2, 256, 15, 20), (30, 40)), ((1, 3, 7, 5), (14, 10)), ((2, 5, 6, 6), (11, 13))):
        x = torch.randn(B, C, h, w, generator=g).cuda()
        cur = torch.randn(B, C, Ht, Wt, generator=g).cuda()
        got = ops.upsample_add(x, cur)
        want = cur + F.interpolate(x, size=(Ht, Wt), mode="bilinear", align_corners=False)
        assert (got - want).abs().max().item() <= 4e-6 * max(1.0, want.abs().max().item())
    # the token-major (channels_last view) map the encoder hands to the FPN step
    xt = torch.randn(2, 15 * 20, 64, generator=g).cuda().transpose(1, 2).reshape(2, 64, 15, 20)
    cur = torch.randn(2, 64, 30, 40, generator=g).cuda()
    want = cur + F.interpolate(xt, size=(30, 40), mode="bilinear", align_corners=False)
    assert (ops.upsample_add(xt, cur) - want).abs().max().item() <= 4e-6 * want.abs().max().item()


@pytest.mark.parametrize("kind,B", [("r50", 4), ("demo", 1)])
def test_three_graphs_in_flight_equal_serial_replay_whole_model(kind, B):
    """The timed configuration of the DEFAULT bench (bench.py --inflight 3): the whole model (cuDNN backbone, head, tail)
    captured three times with separate static buffers and replayed round-robin on three streams. Every concurrent
    replay must reproduce, bit for bit, what the same graph gives when it runs alone - i.e. no kernel of the path keeps
    hidden global scratch that concurrent replays could share, and the derived-weight caches are read-only. (Against the
    EAGER forward only the head is bit-exact, test_graph_replay_equals_eager_config2: cuDNN may choose other TF32
    algorithms under capture.)"""
    from unseenobjectswithmeanshift_b200 import backbones, workloads
    from unseenobjectswithmeanshift_b200.graph import GraphedForward
    backbones.set_tf32(True)
    model = workloads.build_model(kind).cuda()
    inputs = [{k: v.cuda() for k, v in workloads.synthetic_images(kind, B, seed=s, pin=False).items()} for s in range(3)]

    def step(inp):
        outputs, _, padded, _ = model._head_outputs([inp])
        label_map, f = model.label_maps([inp])
        return {"pred_masks": outputs["pred_masks"], "pred_logits": outputs["pred_logits"], "label_map": label_map,
                "scores": f["scores"], "pred_boxes": f["pred_boxes"]}

    with torch.no_grad():
        graphs = [GraphedForward(step, i, warmup=2) for i in inputs]
        alone = []
        for g in graphs:
            g()
            torch.cuda.synchronize()
            alone.append({k: v.clone() for k, v in g.static_out.items()})
        assert not torch.equal(alone[0]["pred_masks"], alone[1]["pred_masks"])
        lanes = [torch.cuda.Stream() for _ in graphs]
        main = torch.cuda.current_stream()
        for rep in range(4):
            for g in graphs:                      # poison the outputs: a replay that did not run would be caught
                for v in g.static_out.values():
                    v.fill_(-1)
            fork = torch.cuda.Event()
            fork.record(main)
            for g, lane in zip(graphs, lanes):
                with torch.cuda.stream(lane):
                    lane.wait_event(fork)
                    g()
            for lane in lanes:
                main.wait_stream(lane)
            torch.cuda.synchronize()
            for g, want in zip(graphs, alone):
                for k in want:
                    assert torch.equal(g.static_out[k], want[k]), (rep, k)
