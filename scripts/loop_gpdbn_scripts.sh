#!/bin/bash
# run the gpdbn_horses reference-script diagnostic several times (intermittent non-finite model_data)
for i in 1 2 3 4 5 6; do python scripts/diag_gpdbn_scripts.py 2>&1 | grep -a "non-finite\|Error\|error\|Traceback" ; done
