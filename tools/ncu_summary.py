"""Print the metrics we track from an .ncu-rep (run here, no GPU needed):
python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'lts__t_sectors.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
def summary(rep):
    rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
        for k in KEYS:
            if k in hdr:
                print('   %-75s %-12s %s' % (k, units[hdr.index(k)], r[hdr.index(k)]))
        stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print('   stall %-60s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))

def traffic(workload, kind, rep, match=None, out="profiles/kernel_traffic.json"):
    """Adds dram__bytes_read.sum + dram__bytes_write.sum of the FIRST kernel in an `ncu --set full` report to
    profiles/kernel_traffic.json under [workload][kind] (bench.py reads `roofline.traffic` from there):
        python tools/ncu_summary.py --traffic cartpole_bnn_b4096 mlp_rollout gpurun_out/prof_r2_mlp_roll.ncu-rep [name part]"""
    import json, os
    rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    r = next(x for x in rows[2:] if match is None or match in x[hdr.index('Kernel Name')])
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot = 0.0
    for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(key)
        tot += float(r[i].replace(',', '')) * scale[units[i]]
    table = json.load(open(out)) if os.path.exists(out) else {}
    table.setdefault(workload, {})[kind] = {
        "dram_bytes_per_launch": tot, "kernel": r[hdr.index('Kernel Name')],
        "gpu_time_us": float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')) * {'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 's': 1e6}[
            units[hdr.index('gpu__time_duration.sum')]],
        "source": "ncu --set full --clock-control none, %s (dram__bytes_read.sum + dram__bytes_write.sum)" % os.path.basename(rep)}
    json.dump(table, open(out, 'w'), indent=1, sort_keys=True)
    print(workload, kind, tot)


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic(*sys.argv[2:6])
    else:
        summary(sys.argv[1])
