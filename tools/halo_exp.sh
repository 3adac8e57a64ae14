for i in 8 0; do
  echo "== case $i"; env PW_HALO_TS=1 timeout -s KILL 60 python tools/umma_probe.py one $i 2>&1 | grep "halo ts" | tail -1 | sed -e 's/.*grid/grid/'
done
timeout -s KILL 120 python tools/umma_probe.py small 2>&1 | sed -e "s/'simt': '[^']*', //" | cut -c1-150 | tail -4
echo "=== default"; timeout -s KILL 200 python tools/umma_probe.py full 2>&1 | sed -e "s/'simt': '[^']*', //" -e "s/'umma': '[^']*', //"
