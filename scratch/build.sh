#!/bin/bash
# rebuild lib/libembrace_sm100.so from any working directory
cd "$(dirname "$0")/.." && python -c "
import embrace_b200
print(embrace_b200.build(force=True))"
