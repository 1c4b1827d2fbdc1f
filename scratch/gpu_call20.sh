#!/bin/bash
T="timeout -k 5"
for m in default default reserve; do $T 120 python scratch/slow_steps.py $m 2>&1 | tail -3; done
