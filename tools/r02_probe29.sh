#!/bin/bash
timeout 900 python -m pytest tests/test_golden.py tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -6
