"""Joins an `ncu --page source --csv` SASS dump of one kernel with `nvdisasm -g` line info of the same cubin and aggregates the
warp-stall samples per source line (innermost line and the outermost inlined-at line in the kernel's own file).
usage: sass_lines.py <ncu source csv> <cubin> <kernel-mangled-substring> [top N]"""
import csv, re, subprocess, sys, collections

def main():
    src_csv, cubin, kname = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    in_k = False
    cur = None            # (innermost, outermost)
    prev_was_loc = False
    lines = {}            # offset -> (inner, outer)
    for l in dis:
        if l.startswith("\t.section") or l.startswith(".section"):
            in_k = (".text." in l) and (kname in l)
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            if prev_was_loc and cur: cur = (cur[0], loc)       # a chain: first line innermost, last line outermost
            else: cur = (loc, loc)
            prev_was_loc = True
            continue
        prev_was_loc = False
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
        if m and cur:
            lines[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {k: i for i, k in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    base = int(body[0][ci["Address"]], 16) if body[0][ci["Address"]].startswith("0x") else int(body[0][ci["Address"]])
    tot = 0
    inner_s, outer_s = collections.Counter(), collections.Counter()
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    outer_st = collections.defaultdict(collections.Counter)
    for r in body:
        a = r[ci["Address"]]
        off = (int(a, 16) if a.startswith("0x") else int(a)) - base
        n = int(r[ci["# Samples"]] or 0)
        tot += n
        inner, outer = lines.get(off, (("?", 0), ("?", 0)))
        inner_s[inner] += n
        outer_s[outer] += n
        for k in stall_cols:
            v = int(r[ci[k]] or 0)
            if v: outer_st[outer][k[6:]] += v
    print(f"total samples {tot}")
    print("== by outermost line (the kernel's own file)")
    for (f, ln), n in outer_s.most_common(topn):
        st = ", ".join(f"{k} {v}" for k, v in outer_st[(f, ln)].most_common(3))
        print(f"  {100.0 * n / tot:5.1f}%  {f}:{ln}   [{st}]")
    print("== by innermost line")
    for (f, ln), n in inner_s.most_common(topn):
        print(f"  {100.0 * n / tot:5.1f}%  {f}:{ln}")

main()
