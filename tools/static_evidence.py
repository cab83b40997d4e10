"""Static evidence for kernels that have not run on a GPU yet: registers / spills from the ptxas logs of
the in-tree build and the SASS instruction mix of libpbx.so -> profiles/r1_static_after_closing.md"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "poissbox_b200", "lib")
WANT = ["fwd_lm_kernel", "bwd_lm_kernel", "periodic_lm_kernel", "k_peer_barrier", "k_peer_allreduce", "k_reduce_peer",
        "k_boundary_thin", "k_update_r", "k_pupdate_x",
        "yz_tma_kernelILb0ELb0ELb0ELb0ELb0ELb0E", "yz_tma_kernelILb1ELb0ELb0ELb0ELb0ELb0E",
        "yz_tma_kernelILb0ELb0ELb0ELb1ELb0ELb0E", "yz_tma_kernelILb1ELb0ELb0ELb1ELb0ELb0E",
        "yz_tma_kernelILb0ELb0ELb0ELb0ELb1ELb0E", "yz_tma_kernelILb1ELb0ELb0ELb0ELb1ELb0E",
        "yz_tma_kernelILb1ELb0ELb0ELb0ELb0ELb1E", "yz_tma_kernelILb1ELb1ELb0ELb0ELb0ELb1E",
        "x_tma_kernelILb1ELb1E", "lineop_yz_tma_kernelILb0ELb0ELb0E", "lineop_yz_tma_kernelILb0ELb0ELb1E",
        "lineop_yz_tma_sum_kernelILb0E", "lineop_x_tma_kernelILb0E", "lineop_x_tma_kernelILb1E",
        "mg_sweep_kernelILi0E", "mg_restrict_kernel", "mg_prolong_kernel"]


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"pbx::\(anonymous namespace\)::", "", r).split("(")[0]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, "libpbx.so")], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur:
            funcs[cur].append(ln)
    regs = {}
    for f in glob.glob(os.path.join(LIB, "*.ptxas.log")):
        txt = open(f).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes spill stores, (\d+) bytes spill loads"
                             r".*?Used (\d+) registers", txt, re.S):
            regs[m.group(1)] = (int(m.group(4)), int(m.group(2)) + int(m.group(3)))
    out = ["# Static evidence for the kernels written after round 1's closing GPU run (no GPU-minutes left)", "",
           "`nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v` (poissbox_b200/csrc/Makefile); registers /",
           "spill bytes from `poissbox_b200/lib/*.ptxas.log`, instruction mix from `cuobjdump -sass poissbox_b200/lib/libpbx.so`",
           "(`python tools/static_evidence.py`). None of these kernels has a measured number yet (DESIGN.md section 10); this file only",
           "shows that they compile to what they are meant to be. The first four `yz_tma_kernel` rows are the measured default kernels",
           "and their swizzled variants, for comparison.", "",
           "| kernel | registers | spill bytes | TMA loads (UTMALDG) | TMA stores (UTMASTG) | mbarrier ops (SYNCS) | sys-scope stores / loads | "
           "FP64 divisions (MUFU.RCP64H) | atomics (ATOMG/RED) |", "|---|---|---|---|---|---|---|---|---|"]
    for w in WANT:
        for name, lines in funcs.items():
            if w in name:
                txt = "\n".join(lines)
                c = lambda pat: len(re.findall(pat, txt))
                r = regs.get(name, ("?", "?"))
                out.append(f"| `{demangle(name)}` | {r[0]} | {r[1]} | {c('UTMALDG')} | {c('UTMASTG')} | {c('SYNCS')} | "
                           f"{c(r'ST[G]?\.E[^;]*STRONG\.SYS')} / {c(r'LD[G]?\.E[^;]*STRONG\.SYS')} | {c(r'MUFU\.RCP64H')} | "
                           f"{c(r'ATOMG|RED\.E')} |")
    out += ["", "Reading: the line-major tridsol kernels and the line operators move all their data by TMA (the UTMALDG counts include the",
            "prologue's copy of each loop's loads) and keep the reference's true divisions (one `MUFU.RCP64H`-seeded Newton sequence per",
            "`__ddiv_rn`); the peer kernels' flag traffic is `ST/LD.E.64.STRONG.SYS` behind `MEMBAR.ALL.SYS`; the swizzled (4th argument),",
            "any-chunk-count (5th) and fused-tail (6th) variants of the y/z kernels stay at the register budget of the measured ones."]
    open(os.path.join(ROOT, "profiles", "r1_static_after_closing.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[8:]))


if __name__ == "__main__":
    main()
