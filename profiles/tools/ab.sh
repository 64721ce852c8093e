# A/B of the two distinct-count paths of fc_agg_finalize on the bench workload.  usage: ab.sh "<modes: part global>" [bench args]
modes="$1"; shift
for m in $modes; do FC_AGG_SETS=$m timeout 300 python bench.py --no-extra --steps 5 --warmup 3 --cpu-sample 2000 "$@" > gpurun_out/ab_$m.log 2> gpurun_out/ab_$m.err; python - <<P
import json
for l in open("gpurun_out/ab_$m.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$m", "$*", round(d["value"]/1e9,3), round(d["ms_per_step"],3), d["detail"]["merge_stages_us"], round(d["detail"]["scan_ms"],3), d.get("parity_checked"), round(d["e2e"]["value"]/1e9,3))
P
tail -2 gpurun_out/ab_$m.err; done
