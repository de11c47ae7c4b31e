"""Print a window of a timeline written by tools/timeline.py: start, end, duration, stream, kernel (tile variant), grid."""
import json, sys
rows = json.load(open(sys.argv[1])); lo = float(sys.argv[2]); hi = float(sys.argv[3])
rows.sort(key=lambda r: r[2])
for r in rows:
    if r[2] < lo or r[2] > hi: continue
    n = r[0]
    tag = 'leaf' if 'leaf' in n else 'cstep' if 'chain' in n else n[n.find('<'):n.find('>') + 1] if 'dgemm' in n else n[:25]
    print(f"{r[2]:8.1f} {r[2]+r[3]:8.1f} {r[3]:7.1f} s{r[1]:<3} {tag} grid {r[4] if len(r) > 4 else '?'}")
leaves = [r for r in rows if 'leaf' in r[0]]
print('span', max(r[2] + r[3] for r in rows), 'leaf periods', [round(b[2] - a[2]) for a, b in zip(leaves[:-1], leaves[1:])], 'last leaf end', leaves[-1][2] + leaves[-1][3])
