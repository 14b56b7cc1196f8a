#!/bin/bash
( timeout 300 python -m pytest tests/test_field_api_gpu.py tests/test_field_gpu.py -x -q ) 2>&1 | tail -4
