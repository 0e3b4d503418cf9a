#!/bin/bash
cd $GRAFT_REPO_ROOT
python scripts/nms_trace.py 1 416 2>&1 | tail -20
python scripts/nms_trace.py 64 608 2>&1 | tail -20
