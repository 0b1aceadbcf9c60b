for v in u1 u2 u4; do python tools/probes/emit_ab.py tools/probes/build/libvoxm_$v.so; done
