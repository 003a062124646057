#!/bin/bash
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print({k: getattr(p, k) for k in dir(p) if 'l2' in k.lower() or 'L2' in k or 'persist' in k.lower() or 'window' in k.lower()})
from cuda import cudart
for name in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrL2CacheSize"):
    a = getattr(cudart.cudaDeviceAttr, name)
    print(name, cudart.cudaDeviceGetAttribute(a, 0))
PY
