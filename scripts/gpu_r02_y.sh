#!/bin/bash
# round 2, call y (1 GPU): programmatic dependent launch on small lattices (CG us per iteration)
timeout 300 python scripts/small_cg_pdl.py 2>&1 | tail -6
