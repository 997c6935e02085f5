"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`):
launch count, total and mean duration, share of the listed launches.  Usage: summarize_launches.py in.csv [out.txt]"""
import collections, csv, sys


def summarize(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * (1.0 if u in ("us", "usecond") else 1e3)
        name = row["Kernel Name"].split("(")[0]
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    out = [f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms in total (per-launch times under ncu are cold-cache and serialised: read the SHARES)"]
    out.append(f"{'kernel':34s} {'launches':>8s} {'total ms':>10s} {'mean us':>9s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append(f"{k:34s} {cnt[k]:8d} {v / 1e3:10.3f} {v / cnt[k]:9.1f} {100 * v / T:6.1f}%")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    text = summarize(sys.argv[1])
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        sys.stdout.write(text)
