for b in 1184 2368 4736 9472; do FLIP_MG_RESTRICT_BLOCKS=$b python scripts/probe_vcycle.py 2>&1 | tail -1 | sed "s/^/restrict_blocks=$b /"; done
