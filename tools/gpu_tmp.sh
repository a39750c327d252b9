#!/bin/bash
timeout 600 python tools/gap_report.py --config tf 2>&1 | grep -v Warning | tail -45
