"""CPU: the tiling policy for non-square / larger inputs (extension)."""
import torch

from textmae_image_compression_b200 import tiling
from textmae_image_compression_b200.distributed import shard_round_robin


def test_kodak_and_div2k_tile_counts():
    assert tiling.tile_grid(512, 768, 224) == (3, 4)            # Kodak landscape: 12 tiles of 224
    assert tiling.tile_grid(768, 512, 224) == (4, 3)            # Kodak portrait
    assert tiling.tile_grid(1080, 2048, 512) == (3, 4)          # DIV2K-shaped: 12 tiles of 512


def test_tile_untile_round_trip_and_zero_padding():
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 512, 768, generator=g)
    tiles = tiling.tile_image(img, 224)
    assert tiles.shape == (12, 3, 224, 224)
    assert torch.equal(tiling.untile_image(tiles, 512, 768), img)
    assert torch.equal(tiles[0], img[:, :224, :224])
    assert torch.equal(tiles[3][:, :, 768 - 3 * 224:], torch.zeros(3, 224, 4 * 224 - 768))     # right padding of the last column
    assert torch.equal(tiles[8][:, 512 - 2 * 224:, :], torch.zeros(3, 3 * 224 - 512, 224))     # bottom padding of the last row
    batch = tiling.tile_image(torch.stack([img, img * 0.5]), 224)
    assert batch.shape == (24, 3, 224, 224) and torch.equal(batch[12:], tiling.tile_image(img * 0.5, 224))


def test_image_rate_from_tiles_and_sharding():
    bpp = torch.full((12,), 0.5)
    total = tiling.image_bpp(bpp, 512, 768, 224)
    assert abs(total.item() - 0.5 * 12 * 224 * 224 / (512 * 768)) < 1e-12
    owned = [shard_round_robin(12, r, 8) for r in range(8)]
    assert sorted(i for o in owned for i in o) == list(range(12)) and max(len(o) for o in owned) == 2
