nproc
for kb in 1024 2048 4096 8192; do for t in 1 4 8 16; do echo "bounce $kb KB, $t threads"; VX_BOUNCE_KB=$kb VX_COPY_THREADS=$t tools/probes/pool_probe batch | grep pageable | cut -c1-150; done; done
