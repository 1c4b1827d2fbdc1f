#!/bin/bash
O=gpurun_out; mkdir -p $O
T="timeout -k 5"
$T 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/c13_tests.log; grep -E "passed|failed|FAILED|assert|Error" $O/c13_tests.log | head -40
