#!/bin/bash
set -u
O=gpurun_out
python tools/track_1m.py 2>&1 | tail -7 | tee $O/r02_tracking_1m_i8_4cell.txt
