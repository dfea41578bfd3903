#!/bin/bash
# Copy the outputs of `profiles/final_run.sh <tag>` (+ optional sanitizer report) from gpurun_out/ into the committed profiles/r2_* files.
# usage (here, after the gpurun call): bash profiles/refresh_artefacts.sh <tag>
set -e
cd "$(dirname "$0")/.."
t=$1; o=gpurun_out
cp $o/${t}_bench.json profiles/r2_bench.json
cp $o/${t}_bench_ref.json profiles/r2_bench_reference_arm.json
cp $o/${t}_configs.json profiles/r2_configs_parity_and_stages.json
cp $o/${t}_launches.csv profiles/r2_launches.csv
cp $o/${t}_sass_opcodes.txt profiles/r2_sass_opcodes.txt
[ -f $o/${t}_sanitizer.txt ] && cp $o/${t}_sanitizer.txt profiles/r2_sanitizer.txt
python profiles/launch_summary.py profiles/r2_launches.csv 14 > profiles/r2_launches_summary.txt
idx() { ncu -i $o/${t}_frame.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; k=h.index('Kernel Name')
names=[r[k] for r in rows[2:]]
want=sys.argv[1:]
out=[]
for w in want:
    for i,n in enumerate(names):
        if n.startswith(w) and i not in out: out.append(i); break
print(' '.join(map(str,out)))" "$@"; }
all=$(ncu -i $o/${t}_frame.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); print(' '.join(str(i) for i in range(len(rows)-2)))")
(echo "# one frame of the headline config, every kernel: ncu --set full --clock-control none (profiles/final_run.sh $t, final build of round 2; times are serialised, cold-cache)"; echo
 for i in $all; do python profiles/ncu_summary.py $o/${t}_frame.ncu-rep $i; echo; done) > profiles/r2_ncu_frame_kernels.txt
# the two tile-sort passes: the onesweep launches with the largest DRAM read
pass=$(ncu -i $o/${t}_frame.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; k=h.index('Kernel Name'); d=h.index('dram__bytes_write.sum'); u=rows[1][d]
c=[(float(r[d])*(1e-3 if u.startswith('K') else 1),i) for i,r in enumerate(rows[2:]) if 'onesweep' in r[k]]
c=sorted(sorted(c)[-2:], key=lambda x:x[1]); print(c[0][1], c[1][1])")
set -- $pass
tail_part=$(sed -n '/^== per-opcode shared-memory wavefronts/,$p' profiles/r2_ncu_onesweep.txt)
(echo "== tile-sort pass 0 (final build of round 2: persistent kernel, 4 look-back chains, half tiles at the end of every segment, atomics ranking, swizzled counters, programmatic dependent launch)"
 python profiles/ncu_summary.py $o/${t}_frame.ncu-rep $1; echo; echo "== tile-sort pass 1"; python profiles/ncu_summary.py $o/${t}_frame.ncu-rep $2; echo; echo "$tail_part") > /tmp/r2_os.txt
cp /tmp/r2_os.txt profiles/r2_ncu_onesweep.txt
python - <<P
import csv, io, json, subprocess
raw = subprocess.run(["ncu", "-i", "$o/${t}_frame.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; r = rows[2 + $1]
def mb(name):
    u = rows[1][h.index(name)]; v = float(r[h.index(name)])
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[u]
d = json.load(open("profiles/onesweep_traffic.json"))
d["dram_bytes_read"] = mb("dram__bytes_read.sum"); d["dram_bytes_write"] = mb("dram__bytes_write.sum")
d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]
json.dump(d, open("profiles/onesweep_traffic.json", "w"), indent=1)
print("onesweep traffic per launch", d["dram_bytes_per_launch"])
P
cat profiles/r2_launches_summary.txt
