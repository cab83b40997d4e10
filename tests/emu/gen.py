"""Source-to-source step of the CPU kernel-logic harness (TEST INFRASTRUCTURE ONLY, see
include/pbx_emu.h): copies poissbox_b200/csrc/*.{cu,cuh,h} to _build/src/ as plain C++,

  * `kernel<<<grid, block, smem, stream>>>(args);`  ->  pbx_emu::launch(grid, block, smem, [&] { kernel(args); });
  * `extern __shared__ T name[];`                   ->  T *name = (T *)pbx_emu::dyn_smem();
  * `__shared__ T name[N];`                         ->  static T name[N];   (one CTA runs at a time)
  * pbx_ptx.cuh (inline PTX)                        ->  include/pbx_ptx_emu.cuh (functional model)

The product sources are not modified and never include anything from here.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "poissbox_b200", "csrc")
OUT = os.path.join(HERE, "_build", "src")


def split_top(s):
    """split on commas that are not nested in (), [], {} or <> of a template argument list"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def match_paren(s, i):
    """index just after the parenthesis matching s[i] == '('"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced parentheses")


LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)


def transform(text, name):
    out, pos, n = "", 0, 0
    while True:
        m = LAUNCH.search(text, pos)
        if not m:
            break
        cfg = split_top(m.group(2))
        if not 2 <= len(cfg) <= 4:
            raise ValueError(f"{name}: cannot parse launch configuration {m.group(2)!r}")
        smem = cfg[2] if len(cfg) > 2 else "0"
        end = match_paren(text, m.end() - 1)
        args = text[m.end():end - 1]
        if text[end:end + 1] != ";":
            raise ValueError(f"{name}: launch not followed by ';'")
        out += text[pos:m.start()]
        out += (f"pbx_emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), "
                f"[&]() {{ {m.group(1)}({args}); }})")
        pos = end
        n += 1
    out += text[pos:]
    out = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
                 r"\1 *\2 = reinterpret_cast<\1 *>(pbx_emu::dyn_smem());", out)
    out = re.sub(r"(?m)^(\s*)__shared__\s+", r"\1static ", out)
    out = out.replace('#include "../../include/pbx.h"', f'#include "{ROOT}/include/pbx.h"')
    if "<<<" in out or "__shared__" in out:
        raise ValueError(f"{name}: untransformed CUDA syntax left")
    return out, n


def main():
    os.makedirs(OUT, exist_ok=True)
    total = 0
    for fn in sorted(os.listdir(SRC)):
        if not fn.endswith((".cu", ".cuh", ".h")):
            continue
        if fn == "pbx_ptx.cuh":
            text = open(os.path.join(HERE, "include", "pbx_ptx_emu.cuh")).read()
            n = 0
        else:
            text, n = transform(open(os.path.join(SRC, fn)).read(), fn)
        dst = os.path.join(OUT, fn[:-3] + ".cpp" if fn.endswith(".cu") else fn)
        if not os.path.exists(dst) or open(dst).read() != text:
            open(dst, "w").write(text)
        total += n
    print(f"gen.py: {total} kernel launches rewritten", file=sys.stderr)


if __name__ == "__main__":
    main()
