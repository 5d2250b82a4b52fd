"""GPU: the mask-select kernel is bit-exact against the verbatim reference routine's committed outputs and against
the C oracle on fresh fuzz (every valid K, ties, empty groups, all-equal, L = 196 and 1024)."""
import pytest
import torch

from oracle import ref_mask
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def test_mask_kernel_bit_exact_on_reference_goldens(cuda_dev, golden_dir):
    cases = torch.load(golden_dir / "mask_golden.pt")
    bad = []
    for c in cases:
        sh, rs, kp = G.mask_select(c["scores"].to(cuda_dev), c["K"], isa=16)
        ref = c["ids_shuffle"]
        ok = torch.equal(sh.cpu(), ref) and torch.equal(rs.cpu(), torch.argsort(ref, dim=1)) and \
            torch.equal(kp.cpu(), ref[:, : c["K"]])
        if not ok:
            bad.append((c["name"], int((sh.cpu() != ref).sum())))
    assert not bad, bad


@pytest.mark.parametrize("isa", [16, 8])
def test_mask_kernel_matches_c_oracle_fuzz(cuda_dev, isa):
    g = torch.Generator().manual_seed(2024 + isa)
    total = mism = 0
    for L, Ks in ((196, (16, 64, 144)), (1024, (64, 256, 576))):
        sc = torch.rand(64, L, generator=g)
        sc[16:32] = torch.round(sc[16:32] * 25) / 25
        sc[32:48] = sc[32:48] ** 4
        sc[48:56] = (torch.randint(0, 40, (8, L), generator=g).float() * torch.randint(0, 165, (8, L), generator=g).float())
        sc[48:56] = sc[48:56] / sc[48:56].amax(dim=1, keepdim=True)
        for K in Ks:
            ref = ref_mask.mask_oracle_c(sc, K, isa=isa)
            sh, _, _ = G.mask_select(sc.to(cuda_dev), K, isa=isa)
            mism += int((sh.cpu() != ref).any(dim=1).sum())
            total += sc.shape[0]
    assert mism == 0, f"{mism}/{total} samples differ"


def test_mask_k_greater_than_l_is_einval(cuda_dev):
    with pytest.raises(ValueError, match="greater than the length"):
        G.mask_select(torch.rand(2, 16, device=cuda_dev), 17)
