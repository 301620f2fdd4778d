"""Summarise an .ncu-rep (read here, no GPU) into the text that is committed under profiles/:
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
Per captured launch: duration, DRAM bytes, pipe utilisation (tensor, XU = MUFU, ALU, FMA, LSU), issue slots, registers."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "shared pipe cycles active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster size"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full summary of {path.split('/')[-1]} (read with `ncu -i ... --page raw --csv`)")
    for r in data:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"\n## launch {d.get('ID')}: {d.get('Kernel Name', '')[:150]}")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"  {label:40s} {d[k]} {u.get(k, '')}")
    det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
    keep = [ln for ln in det.splitlines() if any(t in ln for t in ("Stall", "stall", "highest-utilized", "Issue Slots", "Executed Ipc",
                                                                  "Eligible", "No Eligible", "cycles being stalled"))]
    print("\n## scheduler / stall notes (details page)")
    for ln in keep[:40]:
        print("  " + ln.strip())


if __name__ == "__main__":
    main(sys.argv[1])
