# round 2, call AK: tie-break setter for the EfficientZero / reuse descents; full parity suite
cd $GRAFT_REPO_ROOT
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -n 12 ) | cut -c1-220
