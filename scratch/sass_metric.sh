#!/bin/bash
# developer aid: register count and placement of the 8 LDG.128 of the FJ loop for the exp3 model
make -C /root/repo/gslnls_b200/csrc -s 2>&1 | grep -i "error" -A3
GSLNLS_DUMP_CUBIN=/tmp/e3.cubin python -c "
from gslnls_b200 import Model
m = Model('A * exp(-lam * x) + b', ['A','lam','b'], ['x'], jac=True, fvv=True)
" 2>&1 | grep -A3 "nls_pass'" | grep -E "Used|stack"
cuobjdump -sass -fun nls_pass /tmp/e3.cubin | grep -E "LDG.E.NA.128" | head -8 | awk '{print $1}' | tr '\n' ' '; echo
