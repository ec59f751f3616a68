"""Aggregate an ncu SASS source page by CUDA source line.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv hdl-deflate_b200/libhdlz.so k_compress hdlz_compress
Matches the report's SASS (in order) with `nvdisasm -g` of the cubin extracted from the .so."""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, cubin_tag, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if cubin_tag in f][0]
    txt = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], text=True)
    out, line, infn = [], 0, False
    for l in txt.splitlines():
        if l.startswith(".text."):
            infn = kernel in l
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = int(m.group(2))
            continue
        if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            ins = re.sub(r"/\*.*?\*/", "", l).strip().rstrip(";").strip()
            out.append((line, ins))
    return out


def main():
    sass_csv, so, kernel, tag = sys.argv[1:5]
    src = open(sys.argv[5]).read().splitlines() if len(sys.argv) > 5 else None
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    dis = sass_lines(so, tag, kernel)
    assert len(dis) == len(data), (len(dis), len(data))
    agg = {}
    tot_i = tot_s = 0
    for (line, ins), r in zip(dis, data):
        n = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        wf = int(r[ix["L1 Wavefronts Shared"]] or 0)
        wfi = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        a = agg.setdefault(line, [0, 0, 0, 0, 0])
        a[0] += n; a[1] += s; a[2] += wf; a[3] += wfi; a[4] += 1
        tot_i += n; tot_s += s
    print("total warp-instructions %d, samples %d, SASS instrs %d" % (tot_i, tot_s, len(data)))
    for line in sorted(agg):
        a = agg[line]
        if a[0] * 1000 < tot_i and a[1] * 1000 < tot_s:
            continue
        text = src[line - 1].strip()[:70] if src and 0 < line <= len(src) else ""
        print("%4d inst %5.1f%% samp %5.1f%% smem-wf %11d (ideal %11d) sass %3d | %s" %
              (line, 100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, a[2], a[3], a[4], text))


if __name__ == "__main__":
    main()
