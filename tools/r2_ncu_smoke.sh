mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2zz_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_smoke_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/r2zz_smoke_ncu.log; cut -d, -f5 gpurun_out/r2zz_smoke_launches.csv | sort | uniq -c | sort -rn | head -20
