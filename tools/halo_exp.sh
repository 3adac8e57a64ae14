timeout -s KILL 120 python tools/umma_probe.py small 2>&1 | sed -e "s/'simt': '[^']*', //" | cut -c1-150 | tail -14
timeout -s KILL 200 python tools/umma_probe.py full 2>&1 | sed -e "s/'simt': '[^']*', //" -e "s/'umma': '[^']*', //"
