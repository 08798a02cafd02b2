#!/bin/bash
timeout 300 python scripts/bisect_bf16.py 64 2>&1 | tail -90
