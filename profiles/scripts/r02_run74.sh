#!/bin/bash
VAURA_B200_LIB=$PWD/vaura_b200/_lib/libvaura_b200_rutiming.so python profiles/ru_timing.py 16 2>&1 | tail -14
