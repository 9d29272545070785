#!/usr/bin/env python
"""profiles/r02_ncu_kernels.json from the committed `ncu --set full` summaries (tools/ncu_summary.py output): per bench phase the
DRAM traffic of one launch (dram__bytes_read.sum + dram__bytes_write.sum), the kernel's issue figures and what binds it.
bench.py reads this file for `roofline.traffic` / `binding_resource` (a measured capture of the same kernels, not a run-time value).
usage: python tools/ncu_notes.py > profiles/r02_ncu_kernels.json"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "usecond": 1e-3, "msecond": 1.0}


def load(path):
    rows = list(csv.reader(open(os.path.join(ROOT, "profiles", path))))
    h, u = rows[0], dict(zip(rows[0], rows[1]))
    out = {}
    for r in rows[2:]:
        d = dict(zip(h, r))
        name = d["kernel"].replace("void ", "").split("(")[0]        # keeps the template arguments: psm_factor_kernel<1> != <4>
        if name in out:
            continue                                   # first captured launch of every kernel
        f = lambda k: float(d[k]) if d.get(k) not in (None, "") else None
        out[name] = {
            "ms_under_ncu": f("gpu__time_duration.sum") * UNIT[u["gpu__time_duration.sum"]],
            "dram_bytes_per_launch": f("dram__bytes_read.sum") * UNIT[u["dram__bytes_read.sum"]] + f("dram__bytes_write.sum") * UNIT[u["dram__bytes_write.sum"]],
            "registers": int(f("launch__registers_per_thread")), "grid": int(f("launch__grid_size")), "block": int(f("launch__block_size")),
            "warp_instructions": f("smsp__inst_executed.sum"), "ipc": f("sm__inst_executed.avg.per_cycle_active"),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
            "dram_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "local_ld_sectors": f("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
        }
    return out


BIND = {
    "dsnt_moments_kernel": "HBM: every logit read once (DRAM bytes = algorithmic bytes); issue slots ~52 % busy beside it",
    "psm_factor_kernel": "latency: dependent fp64 mat-vec chains, one CTA per image (one warp per sample in the sequence sampler)",
    "psm_cond_kernel": "latency: per-image conditional maps",
    "sample_contours_kernel": "issue slots (Philox + Box-Muller + 42 x Z mat-vec per thread); 13 warps / SM at 64 000 samples",
    "sample_contours_warp_kernel": "HBM / L2: one 5.7 KB factor streamed per sample",
    "raster_kernel": "issue slots + fixed-latency dependencies: no HBM traffic to speak of (masks live in shared memory); the HBM "
                     "fraction of this kernel is meaningless by construction (SURVEY 8d)",
    "raster_general_kernel": "rare samples only (general flood)",
    "finalize_maps_kernel": "HBM / L2 write: 25 B per pixel out for 4 B in",
    "skew_pdf_grid_kernel": "issue slots (fp64 special functions per grid cell), 2.2 GB written per 512 images",
    "skew_sample_kernel": "issue slots (ALU + XU), tables fed from L2 (hit rate > 93 %)",
    "seq_condition_kernel": "latency / HBM: per-sample 84-dim conditioning",
    "contour_trace_kernel": "DRAM / L2 sectors: one thread walks one mask's boundary, 32 separate sectors per load instruction, the mask read 2.7 times (scan + two walks, evicted in between)",
    "simpson_diameters_kernel": "issue slots: exact integer chord search with exact pruning, one warp per mask",
    "psm_sample_direct_kernel": "latency: the factor recursion on one column per (pair, sample) law, fp64",
    "simpson_volumes_kernel": "latency (small)",
}
PHASES = {
    2: {"dsnt": "dsnt_moments_kernel<1>", "factor": "psm_factor_kernel<4>", "sample": "sample_contours_kernel<34>", "raster": "raster_kernel<0>", "finalize": "finalize_maps_kernel"},
    3: {"dsnt": "dsnt_moments_kernel<1>", "pdf_tables": "skew_pdf_grid_kernel", "skew_sampler": "skew_sample_kernel", "raster": "raster_kernel<0>", "finalize": "finalize_maps_kernel"},
    # the sequence sampler's dominant kernel: the second instant's single sample per law (psm_sample_direct_kernel); config 5's
    # capture (r02h) predates it and still lists the per-sample factor it replaced
    4: {"sequence_sampler": "psm_sample_direct_kernel", "raster_masks": "raster_kernel<0>", "simpson_diameters": "simpson_diameters_kernel", "volumes": "simpson_volumes_kernel"},
    5: {"dsnt": "dsnt_moments_kernel<1>", "sequence_sampler": "psm_factor_kernel<1>", "raster": "raster_kernel<0>", "finalize_fac": "finalize_maps_kernel"},
}
SRC = {2: "r02g_ncu_full_summary_config2_step.csv", 3: "r02h_ncu_full_summary_config3.csv", 4: "r02m_ncu_full_summary_config4.csv", 5: "r02h_ncu_full_summary_config5.csv"}


def main():
    out = {"how": "ncu --set full --clock-control none, first captured launch of each kernel; see the CSVs named under `source`"}
    for c, src in SRC.items():
        k = load(src)
        cfg = {}
        for phase, name in PHASES[c].items():
            if name not in k:
                continue
            e = dict(k[name])
            e.update(kernel=name, source="profiles/" + src, binding=BIND.get(name.split("<")[0]))
            cfg[phase] = e
        cfg["_all_kernels"] = {n: dict(v, binding=BIND.get(n.split("<")[0])) for n, v in k.items()}
        out[f"config{c}"] = cfg
    out["config1"] = {p: dict(v, note="captured at config 2's size (64 x 1000); config 1 is the 1 x 100 latency case of the same kernels") for p, v in out["config2"].items() if not p.startswith("_")}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
