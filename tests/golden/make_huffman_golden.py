"""tests/golden/huffman_refexec.pt: outputs of the reference's HuffmanCoding (utils/huffman.py, executed where it lies) on
`ids_restore`-like inputs.  Run in the build container:  python tests/golden/make_huffman_golden.py"""
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import ref_huffman  # noqa: E402


def cases():
    g = torch.Generator().manual_seed(5)
    out = [("perm196", torch.randperm(196, generator=g)[None]), ("perm1024", torch.randperm(1024, generator=g)[None]),
           ("perm196_batch3", torch.stack([torch.randperm(196, generator=g) for _ in range(3)])),
           ("skewed", (torch.rand(500, generator=g) ** 3 * 40).long()), ("two", torch.tensor([7, 7, 3, 7])),
           ("identity64", torch.arange(64)[None]), ("negatives", torch.randint(-5, 6, (300,), generator=g))]
    mask = torch.load(HERE / "mask_golden.pt")
    ids = mask[0]["ids_shuffle"][:2]                       # real permutations from the verbatim reference routine
    out.append(("kodak_ids_restore", torch.argsort(ids, dim=1)))
    return out


def main():
    Ref = ref_huffman.load_reference_class()
    blob = []
    for name, t in cases():
        h = Ref()
        text, shape, _ = h.compress(t)
        assert torch.equal(h.decompress(text, shape, "cpu"), t)
        blob.append({"name": name, "tensor": t, "text": text, "codes": dict(h.codes)})
    torch.save(blob, HERE / "huffman_refexec.pt")
    print([(b["name"], len(b["text"])) for b in blob])


if __name__ == "__main__":
    main()
