timeout -s KILL 120 python tools/umma_probe.py small 2>&1 | sed -e "s/'simt': '[^']*', //"
timeout -s KILL 200 python tools/umma_probe.py full 2>&1 | sed -e "s/'simt': '[^']*', //"
