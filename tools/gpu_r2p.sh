#!/usr/bin/env bash
set -x
timeout 60 python tools/prof_block.py 20 stamps
