"""The decoder-layer cluster kernel alone (B = 8, Q = 100): timing, or the command ncu wraps:
    ncu --set full --clock-control none --import-source on -k regex:decoder_block -s 2 -c 1 -o gpurun_out/prof python tools/prof_block.py 1"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unseenobjectswithmeanshift_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, Q, C, FF = 8, 100, 256, 2048
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
W = dict(o1=rn(C, C, sc=C ** -0.5), qkv=rn(3 * C, C, sc=C ** -0.5), o2=rn(C, C, sc=C ** -0.5), f1=rn(FF, C, sc=C ** -0.5),
         f2=rn(C, FF, sc=FF ** -0.5), qn=rn(C, C, sc=C ** -0.5), m1=rn(C, C, sc=C ** -0.5), c=rn(3, C, sc=C ** -0.5),
         m2=rn(C, C, sc=C ** -0.5), m3=rn(C, C, sc=C ** -0.5))
bv = {k: rn(v.shape[0], sc=0.1) for k, v in W.items()}
norms = [torch.nn.LayerNorm(C).cuda() for _ in range(4)]
qpos = rn(Q, C)
o, state = rn(B, Q, C), rn(B, Q, C)
with torch.no_grad():
    blob = ops.decoder_block_pack(W["o1"], W["qkv"], W["o2"], W["f1"], W["f2"], W["qn"], W["m1"], W["c"], W["m2"], W["m3"])
    tqk = torch.cat([qpos @ W["qkv"][:2 * C].t(), qpos.new_zeros(Q, C)], 1).contiguous()
    tqn = (qpos @ W["qn"].t()).contiguous()
    bc32 = torch.cat([bv["c"], bv["c"].new_zeros(29)])
    fn = lambda: ops.decoder_block(o, state, blob, b_o1=bv["o1"], norm1=norms[0], b_qkv=bv["qkv"], t_qk=tqk, b_o2=bv["o2"],
                                   norm2=norms[1], b_f1=bv["f1"], b_f2=bv["f2"], norm3=norms[2], block_norm=True,
                                   normd=norms[3], b_qn=bv["qn"], t_qn=tqn, b_m1=bv["m1"], b_c32=bc32, b_m2=bv["m2"],
                                   b_m3=bv["m3"])
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
print(f"decoder_block B{B} Q{Q}: {a.elapsed_time(b) / reps * 1e3:.1f} us per launch")
if len(sys.argv) > 2 and sys.argv[2] == "stamps":
    import ctypes
    from unseenobjectswithmeanshift_b200 import _lib
    L = _lib.lib()
    dbg = torch.zeros(B * 8, 32, dtype=torch.int64, device="cuda")
    L.msmx_decoder_block_debug.argtypes = [ctypes.c_void_p]
    L.msmx_decoder_block_debug(ctypes.c_void_p(dbg.data_ptr()))
    with torch.no_grad():
        fn()
    torch.cuda.synchronize()
    L.msmx_decoder_block_debug(None)
    t = dbg.cpu().double()
    names = ["start", "A(O1) written", "O1 done", "LN1", "gather", "QKV done", "self-attn", "a_free", "gather", "O2 done", "LN2",
             "gather", "F1 done", "h converted", "F2 done", "a_free+scatter", "slabs landed", "LN3+norm+dnorm", "QN done",
             "M1+gather", "M2+gather", "M3 done", "end"]
    d = (t[:, 1:23] - t[:, :22]) / 1.9e3    # us at ~1.9 GHz
    print("stage durations (us): mean over CTAs | max over CTAs")
    for i in range(22):
        print(f"  {names[i + 1]:<18} {d[:, i].mean():7.2f} {d[:, i].max():7.2f}")
    print("  total", ((t[:, 22] - t[:, 0]) / 1.9e3).mean())
