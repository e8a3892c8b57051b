"""Source-level attribution of one profiled kernel launch: where the warp-state samples and the issued instructions of
k_batch_add fall, per code region.

    ncu -i gpurun_out/prof_batch_add_round0.ncu-rep --page source --csv > /tmp/src.csv
    python scripts/ncu_regions.py /tmp/src.csv [top_n]

The report must come from `ncu --set full --import-source on`.  Regions are found from the SASS itself, so the script
survives recompiles: out-of-line subroutines are the targets of CALL.REL (named by their instruction mix: the one with
the most IMAD.WIDE is `mul`, the next `sqr`, the one full of shifts the division-step inverse); the kernel body is cut
at the call of the inverse into "forward loop + warp scan" and "backward loop".  Prints the share of samples / issued
instructions / average active lanes and the top stall reasons per region, then the top_n single instructions by samples.
"""
import csv
import sys


def main(path, top_n=12):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    base = int(data[0][col["Address"]], 16)
    insts = [(int(r[col["Address"]], 16) - base, r[col["Source"]].strip(), r) for r in data]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda r, h: int(r[col[h]] or 0)
    # subroutines = CALL targets
    targets = sorted({int(s.split()[-1].strip(";"), 16) - base for _, s, _ in insts if s.startswith("CALL.REL")})
    bounds = targets + [insts[-1][0] + 16]
    subs = []
    for t, e in zip(targets, bounds[1:]):
        body = [s for a, s, _ in insts if t <= a < e]
        subs.append((t, e, sum("IMAD.WIDE" in s for s in body), sum(s.startswith("SHF.") for s in body)))
    by_wide = sorted(subs, key=lambda x: -x[2])
    names = {}
    if by_wide:
        names[by_wide[0][0]] = "mul"
    if len(by_wide) > 1:
        names[by_wide[1][0]] = "sqr"
    for t, e, w, sh in subs:
        names.setdefault(t, "inverse (division steps)" if sh > 40 else "sub@%#x" % t)
    inv_t = next((t for t, n in names.items() if n.startswith("inverse")), None)
    inv_call = next((a for a, s, _ in insts if inv_t is not None and s.startswith("CALL.REL") and int(s.split()[-1].strip(";"), 16) - base == inv_t), None)

    def region(a):
        for t, e, _, _ in subs:
            if t <= a < e:
                return names[t]
        if inv_call is None:
            return "kernel body"
        return "forward loop + warp scan" if a < inv_call else "backward loop (+ rare paths)"

    tot = {}
    for a, s, r in insts:
        d = tot.setdefault(region(a), {"samples": 0, "inst": 0, "thr": 0})
        d["samples"] += num(r, "# Samples")
        d["inst"] += num(r, "Instructions Executed")
        d["thr"] += num(r, "Thread Instructions Executed")
        for h in stalls:
            d[h] = d.get(h, 0) + num(r, h)
    alls = sum(d["samples"] for d in tot.values()) or 1
    alli = sum(d["inst"] for d in tot.values()) or 1
    print("# %s" % rows[0][1][:150])
    for name, d in sorted(tot.items(), key=lambda kv: -kv[1]["samples"]):
        top = sorted(((v, k[6:]) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:4]
        print("%-30s samples %5.1f%%  inst %5.1f%%  lanes %4.1f  %s" % (
            name, 100 * d["samples"] / alls, 100 * d["inst"] / alli, d["thr"] / max(1, d["inst"]),
            ", ".join("%s %.1f" % (k, 100 * v / max(1, d["samples"])) for v, k in top)))
    print("# top single instructions by samples")
    for n, a, s, r in sorted(((num(r, "# Samples"), a, s, r) for a, s, r in insts), reverse=True)[:top_n]:
        st = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:2]
        print("%5.2f%%  %#07x  %-28s %-56s %s" % (100 * n / alls, a, region(a)[:28], s[:56], ", ".join("%s %d" % (k, v) for v, k in st)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
