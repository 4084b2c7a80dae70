#!/bin/bash
# BASELINE config 5 through the tool itself: interactive_mode on an n=4096, d=10 snapshot, NPTS query points on stdin
# (text "%.17g" and the binary framing), answers on stdout.  Usage: tools/cfg5_stream.sh [NPTS] [DEVICES]
set -e
NPTS=${1:-10000000}
DEVS=${2:-0}
cd "$(dirname "$0")/.."
W=/dev/shm/cfg5
if [ "$(df --output=avail -k /dev/shm | tail -1)" -lt 6000000 ]; then W=/tmp/cfg5; fi
mkdir -p $W; df -h $W | tail -1
gcc -O2 -o $W/gen_points tools/gen_points.c
python tools/cfg5_make_snapshot.py $W/cfg5.snapshot
# generate the points with 8 processes (the generator is printf-bound)
PER=$(( (NPTS + 7) / 8 ))
for i in 0 1 2 3 4 5 6 7; do
  S=$(( i * PER )); C=$PER; if [ $(( S + C )) -gt $NPTS ]; then C=$(( NPTS - S )); fi
  if [ $C -gt 0 ]; then $W/gen_points $S $C 10 > $W/pts.$i.txt & fi
done
wait
cat $W/pts.?.txt > $W/pts.txt; rm -f $W/pts.?.txt
$W/gen_points 0 $NPTS 10 binary > $W/pts.bin
ls -la $W
TOOL=madaiemulator_b200/host/emub_interactive_emulator
export EMUB_STREAM_STATS=1
s=$(date +%s%N); $TOOL interactive_mode $W/cfg5.snapshot --quiet --devices $DEVS < $W/pts.txt > $W/out.txt; e=$(date +%s%N)
echo "text  : $NPTS points in $(( (e - s) / 1000000 )) ms (whole process: CUDA start-up, snapshot load, factorisation included)"
s=$(date +%s%N); cat $W/pts.txt | $TOOL interactive_mode $W/cfg5.snapshot --quiet --devices $DEVS > $W/out_pipe.txt; e=$(date +%s%N)
echo "text through a pipe: $NPTS points in $(( (e - s) / 1000000 )) ms"
cmp $W/out.txt $W/out_pipe.txt && rm -f $W/out_pipe.txt
s=$(date +%s%N); $TOOL interactive_mode $W/cfg5.snapshot --quiet --binary --devices $DEVS < $W/pts.bin > $W/out.bin; e=$(date +%s%N)
echo "binary: $NPTS points in $(( (e - s) / 1000000 )) ms"
s=$(date +%s%N); EMUB_IO_THREADS=1 $TOOL interactive_mode $W/cfg5.snapshot --quiet --devices $DEVS < <(head -n 1000000 $W/pts.txt) > $W/out1.txt; e=$(date +%s%N)
echo "text, EMUB_IO_THREADS=1, first 10^6 points: $(( (e - s) / 1000000 )) ms"
s=$(date +%s%N); $TOOL interactive_mode $W/cfg5.snapshot --quiet --devices $DEVS < /dev/null > /dev/null; e=$(date +%s%N)
echo "start-up only (no points): $(( (e - s) / 1000000 )) ms"
python tools/cfg5_check.py $W $NPTS
rm -rf $W
