#!/bin/bash
echo "--- generic"; Y2_CONV_NO_STREAMK=1 timeout 120 python tools/run_layer.py L14 L19 --iters 10 --raw
echo "--- streamk"; timeout 120 python tools/run_layer.py L14 L19 --iters 10 --raw
