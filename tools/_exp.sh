export PW_HALO_SETS=1
timeout -s KILL 120 python tools/umma_probe.py small 2>&1 | sed -e "s/'simt': '[^']*', //" -e "s/'umma': '[^']*', //" | cut -c1-150 | tail -14
timeout -s KILL 200 python tools/umma_probe.py full 2>&1 | sed -e "s/'simt': '[^']*', //" -e "s/'umma': '[^']*', //"
export PW_HALO_SETS=2
echo SETS2
timeout -s KILL 200 python tools/umma_probe.py full 2>&1 | sed -e "s/'simt': '[^']*', //" -e "s/'umma': '[^']*', //"
