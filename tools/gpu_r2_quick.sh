#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for p in 1 0 1 0; do echo "POOL=$p"; POOL=$p timeout 300 python tools/e2e_loop.py 2>&1 | tail -4 | tr '\n' ' '; echo; done
