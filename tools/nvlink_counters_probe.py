#!/usr/bin/env python
"""Which NVLink byte counters does this box expose?  NVML field values (per link and aggregate), printed before and after a
peer-to-peer copy.  python tools/nvlink_counters_probe.py   (needs >= 2 GPUs)"""
import pynvml as nv
import torch

nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
FIELDS = {"DATA_TX": nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, "DATA_RX": nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX,
          "RAW_TX": nv.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_TX, "RAW_RX": nv.NVML_FI_DEV_NVLINK_THROUGHPUT_RAW_RX}


def read():
    out = {}
    for name, fid in FIELDS.items():
        for scope in [0xFFFFFFFF] + list(range(18)):
            try:
                v = nv.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
                out[(name, scope)] = (v.nvmlReturn, int(v.value.ullVal))
            except Exception as ex:   # noqa: BLE001
                out[(name, scope)] = ("exc", repr(ex)[:60])
    return out


a = read()
x = torch.ones(256 << 20, dtype=torch.uint8, device="cuda:0")
y = torch.empty_like(x, device="cuda:1")
for _ in range(4):
    y.copy_(x)
torch.cuda.synchronize()
b = read()
for k in a:
    if a[k][0] == 0 or b[k][0] == 0:
        print(k, a[k], b[k], "delta KiB", (b[k][1] - a[k][1]) if a[k][0] == 0 and b[k][0] == 0 else None)
print("non-zero return codes:", sorted({v[0] for v in a.values() if v[0] != 0}, key=str))
try:
    for link in range(18):
        print("link", link, "state", nv.nvmlDeviceGetNvLinkState(h, link))
except Exception as ex:  # noqa: BLE001
    print("nvlink state:", repr(ex)[:100])
