#!/bin/bash
# prints the reference vs B200 iteration tables of a bundled driver side by side
exe=$1; shift
cd "$(dirname "$0")/../oracle/_ref"
./$exe "$@" > /tmp/r.txt 2>&1
HIOP_B200=1 ./$exe "$@" > /tmp/b.txt 2>&1
grep -E "^ +[0-9]+ +[-+]?[0-9]\.[0-9]+e" /tmp/r.txt > /tmp/rt.txt
grep -E "^ +[0-9]+ +[-+]?[0-9]\.[0-9]+e" /tmp/b.txt > /tmp/bt.txt
paste -d'|' /tmp/rt.txt /tmp/bt.txt
grep -ci "only" /tmp/r.txt /tmp/b.txt
