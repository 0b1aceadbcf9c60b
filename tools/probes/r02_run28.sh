for z in 4 8; do VX_FILL_ZB=$z python tools/probes/fill_ab.py | cut -c1-200; done
