#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python tools/lnfold_sim.py 2>&1 | tail -6
