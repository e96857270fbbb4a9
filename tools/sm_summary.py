import sys, re
rows = [tuple(map(float, re.findall(r"[\d.]+", l))) for l in sys.stdin if l.startswith("sm ")]
if rows:
    nbmax = max(r[1] for r in rows)
    for name, grp in (("fast", [r for r in rows if r[1] >= 0.5 * nbmax]), ("slow", [r for r in rows if r[1] < 0.5 * nbmax])):
        if grp:
            m = lambda i: sum(r[i] for r in grp) / len(grp)
            print(f"  {name}: {len(grp)} SMs, blocks/SM {m(1):.0f}, warp cycles: total {m(2):.0f} line {m(3):.0f} exact {m(4):.0f} rest {m(2)-m(3)-m(4):.0f}   ids {[int(r[0]) for r in grp][:16]}")
