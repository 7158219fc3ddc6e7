"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals of the LAST step."""
import collections, csv, re, sys

def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        data.append(dict(zip(hdr, r)))
    return data

def main(path, steps=3, full=False):
    data = load(path)
    names = [d["Kernel Name"] for d in data]
    n = len(data) // steps
    for per in range(20, len(names) // 2):     # the launch sequence is periodic: take the period, not a guess
        if names[-per:] == names[-2 * per:-per]:
            n = per
            break
    last = data[-n:]
    agg = collections.OrderedDict()
    seq = []
    for d in last:
        name = re.sub(r"<.*", "", d["Kernel Name"])
        name = re.sub(r"\(.*", "", name)[:70]
        t = float(d["Metric Value"]) / 1000.0
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        seq.append((name, t))
    tot = sum(v[1] for v in agg.values())
    print(f"launches/step {n}  total {tot:.1f} us (serialised, cold cache)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.1f} us {100*v[1]/tot:5.1f}% {v[0]:4d}x  {k}")
    if full:
        for name, t in seq:
            print(f"   {t:8.1f}  {name}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3, "--seq" in sys.argv)
