"""Grouped expert GEMM (dropless MoE) vs a per-expert cuBLAS loop on Mixtral-8x7B shapes (single GPU)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuronx_distributed_training_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    C = ops.C()
    torch.manual_seed(0)
    E, H, I, T, topk = 8, 4096, 14336, 8192, 2
    res = {}
    for name, probs in (("balanced", torch.ones(E)), ("skewed", torch.tensor([8., 4, 2, 1, 1, 0.5, 0.25, 0.25]))):
        idx = torch.multinomial(probs.expand(T, E), topk).reshape(-1).cuda()
        counts = torch.bincount(idx, minlength=E)
        n = T * topk
        dest, be, offs, t_cap = ops.grouped_layout(counts, n)
        x = torch.zeros(t_cap, H, device="cuda", dtype=torch.bfloat16)
        x.index_copy_(0, dest, torch.randn(n, H, device="cuda").bfloat16())
        w1 = (torch.randn(E, H, 2 * I, device="cuda") * 0.02).bfloat16()
        dy = torch.zeros(t_cap, 2 * I, device="cuda", dtype=torch.bfloat16)
        dy.index_copy_(0, dest, torch.randn(n, 2 * I, device="cuda").bfloat16())
        out = torch.empty(t_cap, 2 * I, device="cuda", dtype=torch.bfloat16)
        dx = torch.empty(t_cap, H, device="cuda", dtype=torch.bfloat16)
        dw = torch.zeros(E, H, 2 * I, device="cuda", dtype=torch.float32)
        o = offs.tolist()
        flops = 2.0 * n * H * 2 * I

        def loop_fwd():
            for e in range(E):
                if o[e + 1] > o[e]:
                    torch.matmul(x[o[e]:o[e + 1]], w1[e], out=out[o[e]:o[e + 1]])

        def loop_dgrad():
            for e in range(E):
                if o[e + 1] > o[e]:
                    torch.matmul(dy[o[e]:o[e + 1]], w1[e].t(), out=dx[o[e]:o[e + 1]])

        tmp = torch.empty(H, 2 * I, device="cuda", dtype=torch.bfloat16)

        def loop_wgrad():
            for e in range(E):
                if o[e + 1] > o[e]:
                    torch.matmul(x[o[e]:o[e + 1]].t(), dy[o[e]:o[e + 1]], out=tmp)
                    dw[e].add_(tmp)

        r = {"counts": counts.tolist(), "t_cap": t_cap}
        for tag, native, loop in (
                ("fwd", lambda: C.grouped_gemm(x, w1, out, be, False), loop_fwd),
                ("dgrad", lambda: C.grouped_gemm(dy, w1, dx, be, True), loop_dgrad),
                ("wgrad", lambda: C.grouped_wgrad(x, dy, dw, offs, True), loop_wgrad)):
            tn, tl = timeit(native), timeit(loop)
            r[tag] = {"grouped_ms": round(tn, 3), "loop_cublas_ms": round(tl, 3),
                      "grouped_tflops": round(flops / tn / 1e9, 1), "loop_tflops": round(flops / tl / 1e9, 1)}
            print(f"{name} {tag}: grouped {tn:.3f} ms ({flops / tn / 1e9:.0f} TFLOP/s) vs per-expert cuBLAS loop "
                  f"{tl:.3f} ms ({flops / tl / 1e9:.0f} TFLOP/s)", flush=True)
        # numerics on the forward
        C.grouped_gemm(x, w1, out, be, False)
        ref = torch.cat([x[o[e]:o[e + 1]].float() @ w1[e].float() for e in range(E)])
        err = ((out[:o[-1]].float() - ref).norm() / ref.norm()).item()
        r["fwd_rel_err"] = err
        assert err < 6e-3, err
        res[name] = r
        del w1, dw, out, dy, dx
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/check_grouped.json", "w"), indent=1)
    print("GROUPED_OK")


if __name__ == "__main__":
    main()
