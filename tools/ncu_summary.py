"""Summarise an ncu report (.ncu-rep) into profiles/<name>.json + .txt (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name "free-text note" """
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_throughput_pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_bytes",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed": "utcimma_int8_pct_of_peak",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed": "smem_bank_reads_pct",
    "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed": "smem_bank_writes_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__grid_size": "grid",
    "launch__cluster_size": "cluster_size",
    "launch__registers_per_thread": "registers_per_thread",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_smsp",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9, "Ghz": 1e9,
         "Mhz": 1e6, "Tbyte": 1e12}


def main(rep, out, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS and v != "":
                x = float(v.replace(",", ""))
                d[KEYS[h]] = x * SCALE.get(u, 1.0)
        res.append(d)
    summary = res[0] if len(res) == 1 else {"launches": res}
    summary["note"] = note
    summary["source"] = rep
    json.dump(summary, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        f.write(f"# {note}\n# ncu --set full --clock-control none, report {rep}\n")
        for d in res:
            for k, v in d.items():
                f.write(f"{k:32s} {v}\n")
            f.write("\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
