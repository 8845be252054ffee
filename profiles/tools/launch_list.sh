#!/bin/bash
# usage: scratch/launch_list.sh <out.csv> <command...>   — per-launch device times of every kernel
out=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out" "$@"
