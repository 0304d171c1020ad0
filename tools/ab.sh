#!/bin/bash
# developer tool (GPU box): compare builds of libemc.so placed under variants/ (and the in-tree one)
# usage: tools/ab.sh [lib.so ...]   (EMC_AB_OPTS='{"blocks_per_sm":4,"block_threads":128}' selects launch options)
libs="$@"; [ -z "$libs" ] && libs="erpl_monte_carlo_sim_b200/libemc.so $(ls variants/*.so 2>/dev/null)"
for l in $libs; do EMC_LIB=$PWD/$l python tools/ab_one.py 2>&1 | tail -1; done
