"""NaN-poisoned output check of the single-level tensor-engine splat, for any library build (argv[1])."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from blobctrl_b200 import _capi as C
L = ctypes.CDLL(sys.argv[1])
L.blobsplat_feature_splat.argtypes = C.SIGNATURES["blobsplat_feature_splat"]; L.blobsplat_feature_splat.restype = ctypes.c_int
torch.manual_seed(0)
for (n, k, s, c) in ((3, 33, 64, 320), (64, 33, 64, 320)):
    sc = torch.rand(n, k, s, s, device="cuda"); sc = (sc / sc.sum(1, keepdim=True))
    ft = torch.randn(n, k, c, device="cuda")
    ref = torch.einsum("nkp,nkc->ncp", sc.double().flatten(2), ft.double()).view(n, c, s, s)
    for rep in range(4):
        out = torch.full((n, c, s, s), float("nan"), device="cuda")
        rc = L.blobsplat_feature_splat(sc.data_ptr(), k * s * s, s * s, 1, ft.data_ptr(), out.data_ptr(), n, k, c, s, s, 0, 2, 0,
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        nan = torch.isnan(out)
        err = (out.double() - ref).abs(); err[nan] = 0
        print(f"N={n} rep{rep}: nan count {int(nan.sum())} max err(non-nan) {err.max().item():.3e}")
