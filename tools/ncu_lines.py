#!/usr/bin/env python3
"""Joins an ncu SASS source page (csv) with nvdisasm -g line info: per-source-line share of
executed warp instructions, average active lanes, and stall samples.
usage: ncu_lines.py <sass_page.csv> <nvdisasm_-g.txt> <mangled-kernel-substring> [min_pct]"""
import collections, csv, re, sys, os

def main():
    page, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
    txt = open(dis).read().split('\n')
    i0 = [i for i, l in enumerate(txt) if '.section' in l and kern in l and '.text' in l][0]
    cur, seq = None, []
    for l in txt[i0 + 1:]:
        if '.section' in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            seq.append((int(m.group(1), 16), m.group(2).strip(), cur))
    rows = list(csv.reader(open(page)))
    H = rows[1]
    data = rows[2:2 + len(seq)]
    ii, ti, pi, sm = (H.index(k) for k in ('Instructions Executed', 'Thread Instructions Executed',
                                           'Predicated-On Thread Instructions Executed', '# Samples'))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = 0
    for (off, ins, cur), r in zip(seq, data):
        n = int(r[ii]); tot += n
        a = agg[cur]; a[0] += n; a[1] += int(r[ti]); a[2] += int(r[pi]); a[3] += int(r[sm])
    src = {}
    base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mafrixraytracing_b200', 'csrc')
    for f in os.listdir(base):
        if f.endswith(('.cu', '.cuh', '.h')):
            src[f] = open(os.path.join(base, f)).read().split('\n')
    tots = sum(a[3] for a in agg.values())
    tt = sum(a[1] for a in agg.values()); tp = sum(a[2] for a in agg.values())
    print(f"total warp instr {tot}  avg lanes {tt/tot:.1f}  avg pred-on lanes {tp/tot:.1f}")
    for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or ('', 0))):
        if a[0] < tot * minpct / 100:
            continue
        f, l = key if key else ('?', 0)
        line = src[f][l - 1].strip()[:105] if f in src and l - 1 < len(src[f]) else ''
        print(f"{a[0]/tot*100:5.1f}% lanes={a[1]/max(a[0],1):5.1f} on={a[2]/max(a[0],1):5.1f} stall={a[3]/max(tots,1)*100:5.1f}% {f}:{l:4d} | {line}")

main()
