#!/bin/bash
cd "$(dirname "$0")/.."
TAIL_PROF=1 timeout 300 python tools/tail_ncu.py 2>&1 | tail -40
