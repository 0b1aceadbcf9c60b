cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag
for thp in 0 1; do for t in 8 16; do echo "thp $thp, $t threads"; VX_THP=$thp VX_COPY_THREADS=$t tools/probes/pool_probe batch | grep pageable | cut -c1-150; done; done
