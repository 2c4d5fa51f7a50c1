#!/bin/bash
python profiles/host_glue_profile.py 2>&1 | grep -v "^$" | head -60
