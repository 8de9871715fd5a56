"""Summarise an `ncu --page source --csv` dump: executed warp-instructions and stall samples per contiguous SASS region.
usage: python tools/ncu_src.py file.csv [bucket]   (bucket = number of SASS lines per printed row, default 40)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, iex, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[iex].isdigit():
        if body and r and r[0] == "Kernel Name":
            break                      # first kernel instance only
        continue
    body.append(r)
tot = sum(int(r[iex]) for r in body)
tots = sum(int(r[ism]) for r in body)
print(f"total warp-instructions {tot}  samples {tots}  sass lines {len(body)}")
for b in range(0, len(body), bucket):
    ch = body[b:b + bucket]
    ex = sum(int(r[iex]) for r in ch)
    sm = sum(int(r[ism]) for r in ch)
    ops = {}
    for r in ch:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[iex])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print(f"{b:5d}-{b + len(ch):5d}  instr {ex / tot * 100:5.1f}%  samples {sm / max(tots, 1) * 100:5.1f}%  " +
          " ".join(f"{k}:{v / tot * 100:.1f}" for k, v in top))
