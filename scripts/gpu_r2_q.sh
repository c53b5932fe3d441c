cd "$(dirname "$0")/.."
FIB_PERSIST_TIMELINE=1 timeout 120 python scripts/persist_probe.py 4v 6 2>&1 | grep -A1 timeline | tail -4
FIB_PERSIST_TIMELINE=1 timeout 120 python scripts/persist_probe.py br 6 2>&1 | grep -A1 timeline | tail -4
