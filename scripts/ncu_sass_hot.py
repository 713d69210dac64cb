"""Hot spots of an `ncu --page source --csv` export (SASS view): instructions executed and stall samples per
opcode, and the hottest contiguous address ranges.  usage: ncu_sass_hot.py file.csv [top]"""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ins = []
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        try:
            ins.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ismp])))
        except ValueError:
            pass
    tot = sum(i[2] for i in ins)
    tots = sum(i[3] for i in ins)
    print("instructions executed: %.1f M warp-level, %d stall samples, %d SASS lines" % (tot / 1e6, tots, len(ins)))
    byop = collections.Counter()
    byops = collections.Counter()
    for _, src, ex, smp in ins:
        s = src.split()
        op = s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")
        op = op.split(".")[0]
        byop[op] += ex
        byops[op] += smp
    print("by opcode (share of instructions | share of samples):")
    for op, ex in byop.most_common(18):
        print("  %-10s %6.2f %% | %6.2f %%" % (op, 100.0 * ex / tot, 100.0 * byops[op] / max(tots, 1)))
    # contiguous regions with similar execution count (same basic-block frequency)
    regions = []
    cur = None
    for a, src, ex, smp in ins:
        if cur and ex > 0 and abs(ex - cur[3]) <= 0.02 * max(cur[3], 1):
            cur[1] = a
            cur[2] += ex
            cur[4] += smp
            cur[5] += 1
        else:
            if cur:
                regions.append(cur)
            cur = [a, a, ex, ex, smp, 1, src]
    if cur:
        regions.append(cur)
    regions.sort(key=lambda r: -r[2])
    print("hottest regions (first instr | count per instr | #instr | share instr | share samples):")
    for r in regions[:top]:
        print("  %x  %-44s x%-9d n=%-4d %6.2f %% | %6.2f %%" % (r[0], r[6][:44], r[3], r[5], 100.0 * r[2] / tot, 100.0 * r[4] / max(tots, 1)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
