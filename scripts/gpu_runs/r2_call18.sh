#!/bin/bash
timeout 300 python scripts/exp_two_streams.py 2>&1 | tail -5
