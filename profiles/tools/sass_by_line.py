#!/usr/bin/env python3
"""Join an ncu per-SASS-instruction table with source line info.

usage: sass_by_line.py <report.ncu-rep> <launch-index> <libbedkit.so> <mangled-kernel-substring> [top]

ncu's CSV source page carries per-instruction 'Instructions Executed' and 'Warp Stall Sampling' but no line numbers;
nvdisasm --print-line-info of the same cubin carries the line of every instruction in the same order.  The two are
joined by instruction ordinal inside the kernel, and summed per (file, line)."""
import csv, os, re, subprocess, sys, tempfile, collections


def ncu_sass(rep, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for i, r in enumerate(rows):
        if len(r) > 3 and r[0] == "Address":
            hdr, data = r, rows[i + 1:]
            break
    ii, wi, si = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    ti = hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
    res = []
    for r in data:
        if r and r[0] == "Address":  # ncu repeats the table; the first copy is complete
            break
        if len(r) > max(wi, ii, si) and r[ii].isdigit():
            res.append((r[si], int(r[ii]), int(r[wi]), int(r[ti]) if ti is not None and r[ti].isdigit() else 0))
    return res


def line_table(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in txt:
            continue
        lines, cur, on = [], ("?", 0), False
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                on = kernel in m.group(1)
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur)
        if lines:
            return lines
    raise SystemExit("kernel not found in " + so)


def main():
    rep, launch, so, kernel = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    sass = ncu_sass(rep, launch)
    lines = line_table(so, kernel)
    if len(sass) != len(lines):
        print("warning: %d ncu instructions vs %d nvdisasm instructions" % (len(sass), len(lines)), file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (src, n, st, tn), key in zip(sass, lines):
        a = agg[key]
        a[0] += n; a[1] += st; a[2] += 1
    tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
    print("total warp instructions %d, stall samples %d, sass %d" % (tot, tots, len(sass)))
    print("%-22s %8s %8s %6s" % ("file:line", "inst%", "stall%", "#sass"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %8.2f %8.2f %6d" % ("%s:%d" % key, 100.0 * a[0] / tot, 100.0 * a[1] / max(tots, 1), a[2]))


if __name__ == "__main__":
    main()
