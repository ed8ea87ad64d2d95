"""Which NVLink byte counters does this box expose?  (NVML field values with several scopes,
nvidia-smi nvlink -gt d)  python tools/nvlink_probe.py"""
import subprocess
import pynvml
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("devices", n, "driver", pynvml.nvmlSystemGetDriverVersion())
h = pynvml.nvmlDeviceGetHandleByIndex(0)
F = {"DATA_TX": pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, "DATA_RX": pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX,
     "RAW_TX": pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX, "RAW_RX": pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX}
for scope in (0xFFFFFFFF, 0, 1, 17):
    try:
        v = pynvml.nvmlDeviceGetFieldValues(h, [(f, scope) for f in F.values()])
        print("scope", hex(scope), [(k, x.nvmlReturn, x.valueType, x.value.ullVal) for k, x in zip(F, v)])
    except Exception as e:
        print("scope", hex(scope), "error", e)
try:
    v = pynvml.nvmlDeviceGetFieldValues(h, list(F.values()))
    print("no scope", [(k, x.nvmlReturn, x.valueType, x.value.ullVal) for k, x in zip(F, v)])
except Exception as e:
    print("no scope error", e)
for link in (0, 1):
    try:
        print("link", link, "state", pynvml.nvmlDeviceGetNvLinkState(h, link))
    except Exception as e:
        print("link", link, "state error", e)
for cmd in (["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], ["nvidia-smi", "nvlink", "-s", "-i", "0"]):
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(" ".join(cmd), "rc", r.returncode)
    print(r.stdout[:1500], r.stderr[:300])
