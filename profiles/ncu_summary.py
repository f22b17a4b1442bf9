"""Key metrics of `ncu --set full` captures -> markdown table.

    python profiles/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/<name>.md
Read on the CPU box with `ncu -i <rep> --page raw --csv` (B200_PROFILING.md)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("lts__t_bytes.sum.per_second", "L2 throughput"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [ln for ln in out.splitlines() if not ln.startswith("==")]
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr, units, data = rd[0], rd[1], rd[2:]
    return hdr, units, data


def main():
    for rep in sys.argv[1:]:
        hdr, units, data = rows_of(rep)
        name_i = hdr.index("Kernel Name")
        print(f"### `{rep.split('/')[-1]}`\n")
        for d in data:
            print(f"kernel `{d[name_i][:90]}`\n")
            print("| metric | value |")
            print("|---|---:|")
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"| {label} (`{key}`) | {d[i]} {units[i]} |")
            print()


if __name__ == "__main__":
    main()
