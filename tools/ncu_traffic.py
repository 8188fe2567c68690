"""profiles/ncu_traffic.json + a per-launch table from an `ncu --set full` capture of tools/profile_frame.py:
    python tools/ncu_traffic.py gpurun_out/r2_frame.ncu-rep profiles/r2_frame_kernels.txt profiles/ncu_traffic.json
Every launch of the frame: kernel, grid, dynamic smem, duration, DRAM bytes read + written, L2 hit rate, a few stall figures.
The json maps bench.py's kernel keys to the DRAM bytes per launch (mean over the launches of that kind in the frame)."""
import collections, csv, io, json, re, subprocess, sys

M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
     "launch__shared_mem_per_block_dynamic", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def unit_scale(u):
    u = u.split("/")[0]
    return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main(rep, table, js):
    # a .ncu-rep, or the `ncu -i rep --page raw --csv` dump of one made on the GPU box (the full-frame report is > 64 MiB)
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        g = lambda k: float(r[col[k]].replace(",", "")) * unit_scale(units[col[k]]) if k in col and r[col[k]] not in ("", "n/a") else float("nan")
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        out.append(dict(name=name, grid=grid, smem=g("launch__shared_mem_per_block_dynamic"), us=g("gpu__time_duration.sum"),
                        rd=g("dram__bytes_read.sum"), wr=g("dram__bytes_write.sum"), l2hit=g("lts__t_sector_hit_rate.pct"),
                        l2bytes=g("lts__t_bytes.sum"), warps=g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                        dram_pct=g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        tensor=g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        regs=g("launch__registers_per_thread")))
    # classify the tile-GEMM launches by their grid's role count and position in the layer (QKV: 3 roles; FFN1: 2 roles; out-projection
    # and FFN2: 1 role with the LayerNorm tile in shared memory, alternating; pos MLP / PFN: 1 role, plain shared memory)
    keys = collections.defaultdict(list)
    for o in out:
        k = None
        if "proj_tile_kernel" in o["name"]:
            gy = int(re.findall(r"\d+", o["grid"])[1]) if o["grid"] else 0
            if gy == 8: k = "pos_embed_mlp_x8"
            elif gy == 3: k = "set_attention.qkv_proj_gemm"
            elif gy == 2: k = "ffn_linear1_gelu"
            elif o["smem"] > 90e3:
                k = "set_attention.out_proj_gemm_norm1"       # (the headline frame's only LayerNorm-epilogue tile GEMM)
            else: k = "linear_single_role(pfn / pos_embed_mlp)"
        elif "attn_core_kernel" in o["name"]: k = "set_attention.attn_core"
        elif re.search(r"ffn_fused_kernel<(\(bool\))?(1|true)>", o["name"]): k = "attention_tail_ffn"
        elif "ffn_fused_kernel" in o["name"]: k = "ffn_fused"
        elif "qkv_fused_kernel" in o["name"]: k = "set_attention.qkv_proj_gemm"
        elif "vfe_fused_kernel" in o["name"]: k = "vfe_fused"
        elif "pos_fused_kernel" in o["name"]: k = "pos_embed_tables_x8"
        o["key"] = k or o["name"]
        keys[o["key"]].append(o)
    with open(table, "w") as f:
        f.write(f"# {rep}: every launch of one headline frame (ncu --set full, cold clocks / serialised: shares, not absolutes)\n")
        f.write(f"{'kernel':58s} {'grid':>14s} {'smem':>7s} {'regs':>4s} {'us':>8s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'L2hit%':>6s} {'L2_MB':>8s} {'dram%':>6s} {'warps%':>6s} {'tensor%':>7s}\n")
        for o in out:
            f.write(f"{o['key'][:58]:58s} {o['grid'].replace(' ', ''):>14s} {o['smem']:7.0f} {o['regs']:4.0f} {o['us']:8.1f} {o['rd']/1e6:10.2f} {o['wr']/1e6:10.2f} "
                    f"{o['l2hit']:6.1f} {o['l2bytes']/1e6:8.1f} {o['dram_pct']:6.1f} {o['warps']:6.1f} {o['tensor']:7.1f}\n")
        tot_us = sum(o["us"] for o in out); tot_b = sum(o["rd"] + o["wr"] for o in out)
        f.write(f"\n# frame: {len(out)} launches, {tot_us:.0f} us serialised, {tot_b/1e9:.2f} GB DRAM traffic\n# by kernel kind:\n")
        for k, v in sorted(keys.items(), key=lambda kv: -sum(o["us"] for o in kv[1])):
            us = sum(o["us"] for o in v); b = sum(o["rd"] + o["wr"] for o in v)
            f.write(f"{k[:58]:58s} n={len(v):3d} us={us:8.1f} ({100*us/tot_us:4.1f}%)  dram={b/1e6:8.1f} MB ({100*b/tot_b:4.1f}%)  per launch {b/len(v)/1e6:7.2f} MB\n")
    j = {k: {"dram_bytes_per_launch": round(sum(o["rd"] + o["wr"] for o in v) / len(v)), "launches": len(v),
             "us_per_launch_under_ncu": round(sum(o["us"] for o in v) / len(v), 2),
             "source": f"ncu --set full, {rep.split('/')[-1]} (tools/profile_frame.py), dram__bytes_read.sum + dram__bytes_write.sum"}
         for k, v in keys.items()}
    json.dump(j, open(js, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(*sys.argv[1:4])
