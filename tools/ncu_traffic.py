"""profiles/ncu_traffic.json entry (DRAM bytes and shared-memory wavefronts of one solve-kernel launch) from an `ncu --set full` report:
   python tools/ncu_traffic.py <rep> envs horizon dtype gait source-note"""
import csv, json, subprocess, sys
rep, envs, H, dtype, gait, note = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], sys.argv[6]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = dict(zip(rows[0], zip(rows[2], rows[1])))
def val(k, scale={'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ms': 1e3, 's': 1e6, 'ns': 1e-3}):
    v, u = d[k]
    return float(v) * scale.get(u, 1)
print(json.dumps({"envs": envs, "horizon": H, "dtype": dtype, "gait": gait, "kernel": d['Kernel Name'][0],
                  "dram_bytes_read": int(val('dram__bytes_read.sum')), "dram_bytes_write": int(val('dram__bytes_write.sum')),
                  "smem_wavefronts": int(val('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')),
                  "smem_pipe_pct_of_peak_elapsed_ncu": val('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
                  "l1tex_throughput_pct_active_ncu": val('l1tex__throughput.avg.pct_of_peak_sustained_active'),
                  "kernel_us_under_ncu": val('gpu__time_duration.sum'), "source": note}))
