"""profiles/traffic.json from an ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum launch list of
`bench.py --ncu-mode`: DRAM bytes summed over all launches of each C-ABI entry point in ONE step (last period)."""
import collections, csv, json, re, sys

MAP = {"mlp_bf16_fwd_kernel": "nrc_density_query_fwd", "chain_kernel": "nrc_chain_run", "wgrad_kernel": "nrc_chain_wgrad",
       "encode_bwd_kernel": "nrc_encode_bwd", "mlp_bf16_bwd_kernel": "nrc_density_mlp_bwd",
       "interlevel_loss_kernel": "nrc_interlevel_loss", "density_normals_bwd_kernel": "nrc_density_normals_bwd",
       "encode_fwd_kernel": "nrc_encode_fwd", "grid_regularizer_kernel": "nrc_grid_regularizer_init",
       "encode_tangent_kernel": "nrc_encode_tangent_fwd+bwd"}


def main(path, tag):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    data = [dict(zip(hdr, r)) for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    launches = collections.OrderedDict()
    for d in data:
        launches.setdefault(d["ID"], {"name": d["Kernel Name"], "bytes": 0.0})
        if d["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d["Metric Unit"], 1)
            launches[d["ID"]]["bytes"] += float(d["Metric Value"].replace(",", "")) * mult
    seq = list(launches.values())
    names = [s["name"] for s in seq]
    per = None
    for n in range(20, len(names) // 2 + 1):
        if names[-n:] == names[-2 * n:-n]:
            per = n
            break
    last = seq[-per:] if per else seq
    out, cnt = collections.OrderedDict(), collections.OrderedDict()
    for s in last:
        short = re.sub(r"<.*|\(.*", "", s["name"]).replace("void ", "").replace("nrc::", "").strip()
        ep = MAP.get(short)
        if ep:
            out[ep] = out.get(ep, 0.0) + s["bytes"]
            cnt[ep] = cnt.get(ep, 0) + 1
    out["_launches"] = cnt
    out["_kernel_launches_per_step"] = per
    out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum summed over ALL launches of the entry point in one config-2 "
                    f"step (one period of the CUDA-graph replay; ncu --metrics pass profiles/{tag}_ncu_dram_bytes.csv, cold-cache "
                    f"serialised replays; the --set full capture of the top kernels is profiles/{tag}_ncu_full_top_kernels_summary.txt); bytes")
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
