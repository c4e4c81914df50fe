"""profiles/r02_gemm_dram_traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum
-k regex:k_gemm_bf16 --csv` capture of one step's GEMM launches: average DRAM bytes per launch next
to the algorithmic bytes per launch (A + B + C once each, from bench.py's recorded launch list when
given) — bench.py's `roofline.traffic`.

usage: python tools/gemm_traffic.py gpurun_out/r02_final_gemm_dram.csv [bench.json] > profiles/r02_gemm_dram_traffic.json
"""
import csv
import json
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = defaultdict(dict)
    for r in csv.DictReader(lines):
        per[r["ID"]][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, t_us = 0.0, 0, 0.0
    for _, m in per.items():
        if "dram__bytes_read.sum" not in m:
            continue
        rd, ru = m["dram__bytes_read.sum"]
        wr, wu = m["dram__bytes_write.sum"]
        tot += rd * unit.get(ru, 1.0) + wr * unit.get(wu, 1.0)
        if "gpu__time_duration.sum" in m:
            v, u = m["gpu__time_duration.sum"]
            t_us += v / 1e3 if u in ("ns", "nsecond") else v
        n += 1
    out = {"launches": n, "dram_bytes_per_launch": tot / max(1, n), "dram_bytes_total": tot,
           "kernel_time_us_under_ncu": t_us,
           "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {n} k_gemm_bf16 "
                     f"launches of one eager GPT-2-medium step ({path.split('/')[-1]}; cold-cache, "
                     "serialised launches)"}
    if len(sys.argv) > 2:
        line = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
        roof = line.get("roofline", {})
        out["flops_per_step"] = roof.get("flops_per_step")
    # algorithmic bytes: every operand once.  GPT-2 medium, 8 x 512 tokens: per layer and pass
    # A[4096,K] + B[K,N] + C[4096,N] in bf16 over the 12 GEMMs of a block + LM head
    T, H, V, L = 4096, 1024, 50257, 24
    def g(m, n_, k):
        return 2 * (m * k + k * n_ + m * n_)
    blk = [(T, 3 * H, H), (T, H, H), (T, 4 * H, H), (T, H, 4 * H)]
    alg = 0
    for (m, n_, k) in blk:
        alg += g(m, n_, k)          # forward
        alg += g(m, k, n_)          # dgrad: [T, N] x [N, K]
        alg += g(n_, k, m)          # wgrad: [N, T] x [T, K]
    alg = alg * L + g(T, V, H) + g(T, H, V) + g(V, H, T)
    out["algorithmic_bytes_per_launch"] = alg / (12 * L + 3)
    out["ratio"] = out["dram_bytes_per_launch"] / out["algorithmic_bytes_per_launch"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
