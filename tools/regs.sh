#!/bin/bash
# registers / stack / spills per kernel of the built library (cuobjdump; no GPU needed)
cuobjdump -res-usage 3d-navigation-reinforcement-learning_b200/lib/libnav3d_b200.so 2>/dev/null | paste - - | sed -E 's/_ZN[0-9]+_GLOBAL__N__[0-9a-f_]+nav3d_engine_cu_[0-9a-f]+//' | awk '{print}' | grep -E "${1:-.}" | sed -E 's/Function ([^:]*):/\1/' | cut -c1-200
