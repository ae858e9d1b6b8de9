"""SASS evidence for the hot kernels of libadaprox_cuda.so: per kernel the counts of the mnemonics that prove which hardware path
is used (bulk copies, mbarriers, DSMEM stores, cluster barriers, fp64 tensor-core MMA, ...) and an excerpt of the hot loop.
Runs on the build machine (cuobjdump only, no GPU):   python tools/sass_excerpt.py   ->   profiles/r02_sass_<kernel>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adaptive-proximal-algorithms_b200", "libadaprox_cuda.so")
MNEMONICS = ["UBLKCP", "SYNCS", "STAS", "UCGABAR", "CCTL", "MAPA", "LDS", "STS", "LDGSTS", "LDG", "STG", "DFMA", "DADD", "DMUL", "DMMA", "SHFL",
             "BAR", "ATOM", "RED", "MUFU", "UTMALDG", "UTCMMA", "HMMA", "LDL", "STL", "MEMBAR", "ERRBAR", "NANOSLEEP"]
# kernel (substring of the mangled name) -> (title, regex of the instruction the excerpt is centred on, lines before, lines after)
KERNELS = {
    "k_adapgm_fused": ("single-sweep AdaPGM kernel (solver_fused.cuh): bulk-copy ring, DSMEM exchange", r"STAS", 70, 40),
    "k_adapgm_resident": ("cluster-resident AdaPGM kernel (solver_resident.cuh): DSMEM reduce / broadcast, cluster barriers", r"UCGABAR_ARV", 60, 30),
    "k_adapgm_gridres": ("grid-resident AdaPGM kernel (solver_gridres.cuh): rows in shared memory, partial gradients through L2, two grid barriers", r"LDG\.E\.128\.STRONG\.GPU|LDG\.E\.128", 60, 40),
    "k_primal_dual": ("persistent primal-dual kernel (solver_pd.cuh): bulk-copy ring GEMV phases", r"UBLKCP", 30, 60),
    "k_path_gemm": ("lambda-path contraction (path_gemm.cuh): fp64 tensor-core MMA from a cp.async ring", r"DMMA", 40, 60),
    "k_proxgrad_family": ("backtracking / Nesterov / aGRAAL kernel (solver_pg.cuh)", r"UBLKCP", 20, 40),
    "k_malitsky_pock": ("Malitsky-Pock linesearch kernel (solver_mp.cuh)", r"UBLKCP", 20, 40),
}


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for key, (title, centre, before, after) in KERNELS.items():
        names = [n for n in funcs if key in n]
        if not names:
            print("not found:", key, file=sys.stderr)
            continue
        out = [f"# {title}", f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (arch in the fatbin: {', '.join(arch)})", ""]
        for name in names:
            ins = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in funcs[name] if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
            ops = collections.Counter()
            for l in ins:
                m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
                if m:
                    op = m.group(1)
                    for mn in MNEMONICS:
                        if op.startswith(mn):
                            ops[op if mn in ("SYNCS", "DMMA", "UBLKCP", "STAS", "UCGABAR", "LDS", "LDG", "ATOM", "RED", "CCTL", "LDGSTS") else mn] += 1
                            break
            out.append(f"## {name}: {len(ins)} instructions")
            out.append("mnemonic counts: " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items(), key=lambda kv: (-kv[1], kv[0]))))
            idx = [i for i, l in enumerate(ins) if re.search(centre, l)]
            if idx:
                c = idx[len(idx) // 2]
                out.append(f"excerpt around instruction {c} (matches /{centre}/):")
                out += ["    " + re.sub(r"\s+", " ", l).strip() for l in ins[max(0, c - before): c + after]]
            out.append("")
        path = os.path.join(ROOT, "profiles", f"r02_sass_{key}.txt")
        with open(path, "w") as fh:
            fh.write("\n".join(out) + "\n")
        print(path, sum(len(funcs[n]) for n in names), "lines scanned")


if __name__ == "__main__":
    main()
