#!/usr/bin/env bash
set -x
timeout 300 python tools/dev_timeline.py r50 2>&1 | grep -v Warning | tail -45
timeout 300 python tools/dev_timeline.py r50 aux 2>&1 | grep -v Warning | tail -45
timeout 300 python tools/dev_timeline.py r50-head aux 2>&1 | grep -v Warning | tail -8
