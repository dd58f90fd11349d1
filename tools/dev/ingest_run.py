import sys, importlib.util
spec = importlib.util.spec_from_file_location("t", "tests/test_zz_device_ingest.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
import subprocess
out = subprocess.run([sys.executable, "-c", m.SCRIPT], capture_output=True, text=True, timeout=600)
print("RC", out.returncode); print(out.stdout[-3000:]); print(out.stderr[-6000:])
