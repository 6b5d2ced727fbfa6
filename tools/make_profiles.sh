#!/bin/bash
# Profile captures of round 2 (run under gpurun on ONE GPU; numbers printed by runs under ncu are never bench values).
# Writes gpurun_out/r2_*; tools/ncu_summary.py and the notes in profiles/r2_summary.md turn them into the committed files.
set -x
OUT=gpurun_out
# (1) launch list of the bench command (one span, extras off): every launch of the timed region with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file $OUT/r2_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --spans 1 --no-extras --no-cpu-baseline > $OUT/r2_launches_bench.log 2>&1
# (2) full captures of the step-loop kernels, default path
ncu --set full --clock-control none --import-source on -k regex:"k_time|k_freq" -s 40 -c 6 -o $OUT/r2_full_ssfm \
    python tools/prof_ssfm.py --steps 12 > $OUT/r2_full_ssfm.log 2>&1
# (3) the tensor-map frequency pass and the bulk-copy-fed time pass (options)
OCB_FREQ_TMA=1 OCB_TIME_KERNEL=bulk ncu --set full --clock-control none --import-source on -k regex:"k_time_bulk|k_freq_tma" -s 40 -c 6 \
    -o $OUT/r2_full_ssfm_tma python tools/prof_ssfm.py --steps 12 > $OUT/r2_full_ssfm_tma.log 2>&1
# (4) receiver chain: launch list + full capture of bps, equalizer and the EDC kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_rx.csv python tools/prof_rx.py 19 > $OUT/r2_launches_rx.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bps|k_mimo_eq_la|k_edc" -c 6 -o $OUT/r2_full_rx \
    python tools/prof_rx.py 17 > $OUT/r2_full_rx.log 2>&1
for f in r2_full_ssfm r2_full_ssfm_tma r2_full_rx; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/$f.raw.csv > $OUT/$f.summary.txt
  cat $OUT/$f.summary.txt
  rm -f $OUT/$f.ncu-rep   # gpurun brings back at most 64 MiB: the raw metric table is what gets committed
done
# (5) second session: warm-L2 DRAM traffic of the step loop (single pass, caches left alone), units-in-flight sweep,
#     transmitter kernels
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:"k_time|k_freq" -s 100 -c 40 --csv --log-file $OUT/r2_inloop_traffic.csv python tools/prof_ssfm.py --steps 40 > $OUT/r2_inloop_traffic.log 2>&1
python tools/unit_concurrency.py 16 2>&1 | tail -8 | tee $OUT/r2_unit_concurrency.jsonl
ncu --set full --clock-control none -k regex:"k_wdm_combine|k_iqm_power|k_row_absmax2|k_upsample" -s 4 -c 4 -o $OUT/r2_full_tx \
    python tools/prof_tx.py > $OUT/r2_full_tx.log 2>&1
ncu -i $OUT/r2_full_tx.ncu-rep --page raw --csv > $OUT/r2_full_tx.raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/r2_full_tx.raw.csv > $OUT/r2_full_tx.summary.txt
rm -f $OUT/r2_full_tx.ncu-rep
