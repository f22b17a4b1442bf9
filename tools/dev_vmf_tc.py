"""GPU dev check of the tcgen05 vMF attention kernel: error vs an fp64 reference and timing vs the
CUDA-core kernel. Each configuration runs in a fresh process (env switches are read once) under a timeout."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref64(q, k, v, blocked, kappa, nq, nk):
    q, k, v = q.double(), k.double(), v.double()
    if nq: q = torch.nn.functional.normalize(q, dim=-1, eps=1e-12)
    if nk: k = torch.nn.functional.normalize(k, dim=-1, eps=1e-12)
    s = kappa * q @ k.transpose(-1, -2)
    if blocked is not None:
        s = s.masked_fill(blocked, float("-inf"))
    o = torch.softmax(s, -1) @ v
    return torch.nn.functional.normalize(o, dim=-1, eps=1e-12)


def child(quick):
    global torch
    import torch
    from unseenobjectswithmeanshift_b200 import ops
    dev = torch.device("cuda")
    cases = [  # B, H, Q, S, hd, masked, shared(k==v, no normalisation)
        (1, 1, 100, 128, 32, False, False), (1, 2, 100, 300, 32, True, False), (2, 8, 100, 1200, 32, True, False),
        (8, 8, 100, 4800, 32, True, False), (8, 8, 100, 100, 32, False, False), (1, 1, 100, 5000, 64, False, True),
        (2, 1, 100, 777, 64, False, False), (1, 8, 37, 1000, 32, True, False)]
    if not quick:
        cases += [(1, 8, 100, 307200, 32, True, False), (4, 1, 100, 307200, 64, False, True)]
    for (B, H, Q, S, hd, masked, shared) in cases:
        g = torch.Generator(device="cuda").manual_seed(S + hd)
        C = H * hd
        q = torch.randn(B, Q, C, device=dev, generator=g)
        k = torch.randn(B, S, C, device=dev, generator=g)
        v = k if shared else torch.randn(B, S, C, device=dev, generator=g)
        kappa = 10.0 if shared else 30.0
        if shared:
            k = v = torch.nn.functional.normalize(k, dim=-1)
            q = torch.nn.functional.normalize(q, dim=-1)
        hv = lambda t: t.unflatten(-1, (H, hd)).permute(0, 2, 1, 3)
        bits = ro = blocked = None
        if masked:
            blocked = torch.rand(B, Q, S, device=dev, generator=g) < 0.5
            blocked[:, 3] = True  # a row that blocks everything -> treated as open
            words = (S + 31) // 32
            pad = torch.zeros(B, Q, words * 32, dtype=torch.bool, device=dev)
            pad[..., :S] = blocked
            sh = torch.arange(32, device=dev, dtype=torch.int64)
            bits = (pad.view(B, Q, words, 32).long() << sh).sum(-1)
            bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32).contiguous()
            ro = (~blocked).any(-1).to(torch.int32).contiguous()
            blocked = blocked & (ro != 0).unsqueeze(-1)
        out = ops.vmf_attention(hv(q), hv(k), hv(v), blocked_bits=bits, row_open=ro, kappa=kappa,
                                normalize_q=not shared, normalize_k=not shared)
        torch.cuda.synchronize()
        err = float("nan")
        if B * H * Q * S <= 4e8:
            r = ref64(hv(q), hv(k), hv(v), None if blocked is None else blocked.unsqueeze(1), kappa, not shared, not shared)
            err = (out.double() - r).abs().max().item()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5 if S > 100000 else 20
        for _ in range(3): ops.vmf_attention(hv(q), hv(k), hv(v), blocked_bits=bits, row_open=ro, kappa=kappa, normalize_q=not shared, normalize_k=not shared)
        a.record()
        for _ in range(n): ops.vmf_attention(hv(q), hv(k), hv(v), blocked_bits=bits, row_open=ro, kappa=kappa, normalize_q=not shared, normalize_k=not shared)
        b_.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b_) / n * 1e3
        fl = 4.0 * B * H * Q * S * hd
        by = 4.0 * B * S * C * (1 if shared else 2)
        print(f"  B{B} H{H} Q{Q} S{S} hd{hd} mask{int(masked)} shared{int(shared)}: max abs err {err:.3e}  {us:9.1f} us "
              f"{fl / us / 1e6:8.2f} TFLOP/s {by / us / 1e3:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(len(sys.argv) > 2 and sys.argv[2] == "quick")
    else:
        quick = "quick" if (len(sys.argv) > 1 and sys.argv[1] == "quick") else "full"
        for env in ({"MSM_DISABLE_TC": "1"}, {}):
            print("env", env, flush=True)
            e = dict(os.environ); e.update(env)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", quick], env=e, timeout=150)
                print("  rc", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("  TIMEOUT (kernel hang?)", flush=True)
