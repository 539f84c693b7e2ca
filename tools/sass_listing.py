"""Committed SASS evidence (north_star: "each choice is evidenced by ncu ... with a committed SASS listing").

    python tools/sass_listing.py        # runs here (no GPU needed): cuobjdump -sass on the built library

Writes profiles/sass/<family>.txt: for every kernel of the family the tensor / TMA / TMEM instruction census and the
instruction lines themselves (address + mnemonic + operands) for
  UTCHMMA (tcgen05.mma) | LDTM/STTM (tcgen05.ld/st) | UTMALDG/UTMASTG/UTMAREDG (cp.async.bulk.tensor) |
  UBLKCP (cp.async.bulk) | UTCBAR (tcgen05.commit) | HMMA (legacy mma.sync) | LDGSTS (cp.async) | REDG/ATOMG.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multilingual-image-captioning_b200", "libmic_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF", "HMMA", "LDGSTS",
        "LDSM", "SYNCS", "REDG", "ATOMG", "MUFU.EX2", "SHFL"]
FAMILIES = [("gemm_tcgen05", r"gemm_kernel"), ("decoder_step", r"decoder_step_kernel|pack_weight_tiles|pack_cross_kv|barrier_bench"),
            ("attention", r"attention"), ("beam_search", r"beam_|greedy_|search_merge"),
            ("elementwise", r".*")]
MAX_LINES = 6       # instruction lines kept per mnemonic per kernel


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.split("\n")


def main():
    cmd = ["cuobjdump", "-sass", LIB]
    txt = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            funcs[cur].append(line.strip())
    names = list(funcs)
    pretty = dict(zip(names, demangle(names)))
    os.makedirs(OUT, exist_ok=True)
    taken = set()
    for fam, pat in FAMILIES:
        rows = [n for n in names if n not in taken and re.search(pat, pretty[n])]
        taken.update(rows)
        with open(os.path.join(OUT, fam + ".txt"), "w") as f:
            f.write(f"# {' '.join(cmd[:2])} multilingual-image-captioning_b200/libmic_b200.so   (sm_100a, nvcc 12.9)\n")
            f.write(f"# family '{fam}': {len(rows)} kernels.  Census = instruction counts; then up to {MAX_LINES} lines per mnemonic.\n\n")
            for n in rows:
                ins = funcs[n]
                cnt = collections.Counter()
                keep = collections.defaultdict(list)
                for l in ins:
                    body = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip()
                    op = body.split()[0] if body.split() else ""
                    if op.startswith("@"):
                        op = body.split()[1] if len(body.split()) > 1 else ""
                    for k in KEYS:
                        if op.startswith(k):
                            cnt[k] += 1
                            if len(keep[k]) < MAX_LINES:
                                keep[k].append(l[:150])
                f.write(f"== {pretty[n][:230]}\n   {len(ins)} SASS instructions; " +
                        ", ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k]) + "\n")
                for k in KEYS:
                    for l in keep[k]:
                        f.write("      " + l + "\n")
                f.write("\n")
        print(fam, len(rows), "kernels")


if __name__ == "__main__":
    main()
