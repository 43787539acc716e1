#!/usr/bin/env python3
"""Per-region view of an ncu source page: blocks of N SASS instructions with their share of samples, executed
instructions per unit, dominant opcodes and stall reasons.  Usage: ncu_regions.py <source.csv | .ncu-rep> [units] [block]"""
import collections, csv, io, subprocess, sys

def main(path, units=1.0, blk=60):
    units, blk = float(units), int(blk)
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(text)))
    else:
        rows = list(csv.reader(open(path)))
    h = rows[1]; ix = {n: i for i, n in enumerate(h)}
    data = rows[2:]
    def f(r, n):
        try: return float(r[ix[n]])
        except Exception: return 0.0
    tot = sum(f(r, '# Samples') for r in data) or 1
    for i in range(0, len(data), blk):
        chunk = data[i:i + blk]
        s = sum(f(r, '# Samples') for r in chunk); ie = sum(f(r, 'Instructions Executed') for r in chunk)
        ops = collections.Counter(); st = collections.Counter()
        for r in chunk:
            t = r[ix['Source']].split(); ops[t[1] if t[0].startswith('@') else t[0]] += 1
            for n in h:
                if n.startswith('stall_') and '(Not' not in n: st[n] += f(r, n)
        top = ', '.join(f"{k[6:]}:{v / max(s, 1):.2f}" for k, v in st.most_common(3))
        print(f"{i:5d} samp={s / tot:.3f} inst/unit={ie / units:8.1f} ops={dict(ops.most_common(4))} | {top}")

if __name__ == "__main__":
    main(*sys.argv[1:])
