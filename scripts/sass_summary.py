"""Instruction counts per kernel of the built library for the mnemonics that show which hardware paths are used
(python scripts/sass_summary.py > profiles/r2_sass_summary.txt; needs cuobjdump, no GPU)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UBLKCP", "UTMACMDFL", "UTMALDG", "UTMASTG", "ACQBULK", "ATOMS", "ATOMG", "RED.", "REDUX", "MATCH", "SHFL", "I2F.S64", "IMAD.WIDE", "DADD", "MUFU.RCP", "HMMA", "UTCMMA"]
HEAD = """# cuobjdump -sass globalillumination_b200/libshadowgi.so (sm_100a cubin, final build of round 2): instruction counts per kernel for the
# mnemonics that show which hardware paths are used.  UBLKCP = cp.async.bulk (TMA engine, linear form: depth-tile flush),
# UTMACMDFL(USH) = bulk-group commit, UTMALDG / UTMASTG = tensor-map TMA (not used: see DESIGN.md 5a), ACQBULK = griddepcontrol.wait
# (programmatic dependent launch), ATOMS = shared-memory atomics (tile payload), ATOMG / RED = global atomics (tile cursors, stencil
# counts of list segments), REDUX = warp reductions, MATCH = __match_any_sync (aggregated atomics), I2F.S64 = 64-bit edge value ->
# float, IMAD.WIDE = 32x32->64 edge products and steps, DADD = exact double stepping of the region-covering sweep, MUFU.RCP = reciprocals
# (k_visibility_multi*: one per projective divide + the plain-division fallback).  No tensor-core instructions (HMMA / UTCMMA):
# nothing on this path is a dense contraction.
"""


def main():
    lib = os.path.join(ROOT, "globalillumination_b200", "libshadowgi.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    rows, name, cnt, n = [], None, None, 0
    def flush():
        if name:
            rows.append((name, n, cnt))
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            full = m.group(1)
            k = re.search(r"(k_[A-Za-z0-9_]+?)(E[vN]|ENS_|$)", re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f_]+?(sgi_\w+?_cu)_[0-9a-f]+\d\d", "", full))
            name, cnt, n = (k.group(1) if k else full)[:64], {c: 0 for c in COLS}, 0
            continue
        if name and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            n += 1
            for c in COLS:
                if re.search(r"\b" + re.escape(c), line):
                    cnt[c] += 1
    flush()
    print(HEAD)
    print(f"{'kernel':64s} {'instr':>7s} " + " ".join(f"{c:>9s}" for c in COLS))
    for name, n, cnt in rows:
        print(f"{name:64s} {n:7d} " + " ".join(f"{cnt[c]:9d}" for c in COLS))


if __name__ == "__main__":
    main()
