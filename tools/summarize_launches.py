"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: device time, share and launch count per
kernel (names shortened to the part before the argument list).  usage: summarize_launches.py <csv> [top]"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    return name[:110]


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    t, n = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        t[k] += float(r[vi].replace(",", "")) * 1e-6
        n[k] += 1
    total = sum(t.values())
    print(f"total device time: {total:.1f} ms over {sum(n.values())} launches\n")
    print("| ms | share | launches | kernel |\n|---|---|---|---|")
    for k in sorted(t, key=t.get, reverse=True)[:top]:
        print(f"| {t[k]:.3f} | {100 * t[k] / total:.1f}% | {n[k]} | `{k}` |")
    own = ("gemm_tma_ws_kernel", "gemm_kernel", "transform_kernel", "relayout_kernel", "permute_kernel", "dot_kernel", "axpy",
           "scale_kernel", "lincomb_kernel", "splitk_reduce", "stage_", "peer_allreduce", "gather_cols", "scatter_cols", "tnl::")
    ours = sum(v for k, v in t.items() if any(o in k for o in own))
    gemm = sum(v for k, v in t.items() if "gemm_tma_ws_kernel" in k or "gemm_kernel" in k)
    print(f"\nkernels of this library (tnl::*): {ours:.1f} ms = {100 * ours / total:.1f}% ; grouped DGEMM: {gemm:.1f} ms = "
          f"{100 * gemm / total:.1f}%")


if __name__ == "__main__":
    main()
