mkdir -p gpurun_out
for c in "2 0 1" "1 0 0" "1 2 0"; do
  set -- $c
  python tools/fixture_case_launches.py $1 $2 $3 2>&1 | grep call
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2n_case_$1$2$3.csv python tools/fixture_case_launches.py $1 $2 $3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2n_case_$1$2$3.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]:
    print("   ", r[4][:64].ljust(64), r[8], r[-1])
PY
done
