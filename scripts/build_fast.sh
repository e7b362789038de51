#!/bin/bash
# development build: only the kernel variants the benchmark codes need (-DQD_FAST_BUILD)
python __graft_entry__.py --force --fast "$@"
