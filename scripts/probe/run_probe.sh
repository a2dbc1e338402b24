#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}/scripts/probe"
run() { echo -n "$* -> "; ./tma_probe "$@" 2>&1 | tail -1; }
# width pitch D C bx by bz bc x0 y0 z0
run 35 36 35 8 36 19 7 2 0 0 0      # conv2 fwd (works in the kernel)
run 32 32 32 8 40 15 7 2 -3 -3 -3   # conv2 dgrad (fails)
run 32 32 32 8 40 15 7 2 -4 -3 -3   # aligned negative x
run 32 32 32 8 40 15 7 2 0 -3 -3    # no negative x
run 32 32 32 8 40 15 7 2 0 0 0      # box wider than tensor only
run 32 32 32 8 32 15 7 2 0 -3 -3    # negative y,z only
run 32 32 32 8 32 15 7 2 1 0 0      # unaligned positive x
run 32 32 32 8 36 15 7 2 -3 0 0
run 32 32 32 8 40 12 7 2 -3 -3 -3
run 32 32 32 8 40 16 7 2 -3 -3 -3
run 32 32 32 8 40 15 4 2 -3 -3 -3
run 16 16 16 8 24 13 7 2 -3 -3 -3   # conv1 dgrad
