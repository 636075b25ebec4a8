"""
Golden-fixture case table shared by make_golden.py (which runs the REAL
reference from /root/reference in the build container) and the tests (which
check oracle/ and the CUDA path against the committed fixtures).
"""

TINY = dict(depth=4, dim=32, heads=2, mlp_ratio=4, position_encoding_size=(4, 4),
            window_indices=(0, 2), window_size=(4, 4), relative_embedding_size=(5, 5))
TINY_GLOBAL = dict(depth=2, dim=32, heads=2, mlp_ratio=4, position_encoding_size=(4, 4))
SMALL_B = dict(depth=3, dim=768, heads=12, mlp_ratio=4, position_encoding_size=(14, 14),
               window_indices=(0, 2), window_size=(14, 14), relative_embedding_size=(64, 64))

# no rel-pos: the reference's interpolation path (utils.py:179) only works for square global grids
SMALL_TC = dict(depth=3, dim=768, heads=12, mlp_ratio=4, position_encoding_size=(14, 14),
                window_indices=(0, 2), window_size=(14, 14))

# the full ViTDet-B backbone (12 blocks, 8 windowed + 4 global) at BASELINE.json configs[0]: 672 x 672 -> 42 x 42 tokens,
# nine 14 x 14 windows without padding, global rel-pos tables interpolated from 64 to 42 entries, k = 512 of 1764
VITDET_B_FULL = dict(depth=12, dim=768, heads=12, mlp_ratio=4, position_encoding_size=(14, 14),
                     window_indices=(0, 1, 3, 4, 6, 7, 9, 10), window_size=(14, 14), relative_embedding_size=(64, 64))

# global blocks with interpolated rel-pos tables on a 6 x 6 grid: K / V pooled 2 x 2 -> 9 keys (oracle only)
TINY_REL = dict(depth=2, dim=32, heads=2, mlp_ratio=4, position_encoding_size=(4, 4), relative_embedding_size=(5, 5))

CASES = {
    # windowed (padded 7->8) + global eventful blocks, rel-pos with interpolated tables
    "tiny_vitdet": dict(cfg=TINY, input_size=(7, 7), batch=1, frames=5, policy=("topk", dict(k=12)),
                        block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                        std=0.08, stream="drift", seed=1),
    # batch > 1, class token, no windows / rel-pos (the ViViT spatial sub-model shape)
    "tiny_vivit": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=3, frames=4, policy=("topk", dict(k=5)),
                       block_class="EventfulBlock", has_class_token=True, std=0.08, stream="drift", seed=2),
    "tiny_matmul1": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=2, frames=4, policy=("topk", dict(k=6)),
                         block_class="EventfulMatmul1Block", std=0.08, stream="drift", seed=3),
    "tiny_tokenwise": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=2, frames=4, policy=("topk", dict(k=6)),
                           block_class="EventfulTokenwiseBlock", std=0.08, stream="drift", seed=4),
    "tiny_threshold": dict(cfg=TINY, input_size=(7, 7), batch=1, frames=4,
                           policy=("threshold", dict(threshold=1.0)),
                           block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                           std=0.08, stream="patch", seed=5),
    "tiny_fraction": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=1, frames=3,
                          policy=("fraction", dict(fraction=0.4)),
                          block_class="EventfulBlock", std=0.08, stream="drift", seed=6),
    "tiny_cast": dict(cfg=TINY, input_size=(7, 7), batch=1, frames=4, policy=("topk", dict(k=12)),
                      block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                      matmul_2_cast="bfloat16", std=0.08, stream="drift", seed=7),
    "tiny_gate_before_ln": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=1, frames=4,
                                policy=("topk", dict(k=6)), block_class="EventfulBlock",
                                gate_before_ln=True, std=0.08, stream="drift", seed=8),
    "tiny_stgt": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=1, frames=4, policy=("topk", dict(k=6)),
                      block_class="EventfulTokenwiseBlock", stgt=True, std=0.08, stream="drift", seed=9),
    "tiny_dense": dict(cfg=TINY, input_size=(7, 7), batch=2, frames=2, policy=None,
                       block_class="Block", windowed_class=None, std=0.08, stream="drift", seed=10),
    # real ViTDet-B widths (D=768, H=12, 14x14 windows padded 16->28, 64-entry rel tables interpolated)
    "small_vitdet_b": dict(cfg=SMALL_B, input_size=(16, 16), batch=1, frames=3, policy=("topk", dict(k=64)),
                           block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                           std=0.04, stream="drift", seed=11, subsample=True),
    # 64-wide token grid, N = 512: the global block takes the tcgen05 attention kernels
    "small_vitdet_tc": dict(cfg=SMALL_TC, input_size=(8, 64), batch=2, frames=3, policy=("topk", dict(k=160)),
                            block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                            std=0.04, stream="drift", seed=12, subsample=True),
    "tiny_pool_dense": dict(cfg=TINY_REL, input_size=(6, 6), batch=2, frames=2, policy=None, block_class="Block",
                            windowed_class=None, pool_size=(2, 2), std=0.08, stream="drift", seed=14),
    "tiny_pool_matmul1": dict(cfg=TINY_REL, input_size=(6, 6), batch=1, frames=4, policy=("topk", dict(k=10)),
                              block_class="EventfulMatmul1Block", pool_size=(2, 2), std=0.08, stream="drift", seed=15),
    "tiny_pool_eventful": dict(cfg=TINY_REL, input_size=(6, 6), batch=1, frames=4, policy=("topk", dict(k=10)),
                               block_class="EventfulBlock", pool_size=(2, 2), std=0.08, stream="drift", seed=16),
    "vitdet_b_672": dict(cfg=VITDET_B_FULL, input_size=(42, 42), batch=1, frames=3, policy=("topk", dict(k=512)),
                         block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                         std=0.02, stream="drift", seed=13, subsample=True),
    # the reference's timed CUDA configuration (configs/time/vitdet_vid/_cuda.yml): fp32 model, fp16 attention-value path
    "small_cast16": dict(cfg=SMALL_B, input_size=(16, 16), batch=1, frames=4, policy=("topk", dict(k=64)),
                         block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock", matmul_2_cast="float16",
                         std=0.04, stream="drift", seed=18, subsample=True),
    # the "spatial" configuration (configs/evaluate/vitdet_vid/_spatial.yml): 2 x 2 K/V pooling on the global block only
    "small_pool": dict(cfg=SMALL_B, input_size=(16, 16), batch=1, frames=4, policy=("topk", dict(k=64)),
                       block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock", pool_size=(2, 2),
                       std=0.04, stream="drift", seed=19, subsample=True),
    # the BENCHMARKED configuration (BASELINE configs[1]): 1024 x 1024 -> 64 x 64 tokens, 25 padded 14 x 14 windows,
    # k = 2048 of 4096; 4 frames so that the CUDA-graph path (captured at frame 2) is replayed at least once
    "vitdet_b_1024": dict(cfg=VITDET_B_FULL, input_size=(64, 64), batch=1, frames=4, policy=("topk", dict(k=2048)),
                          block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                          std=0.02, stream="drift", seed=17, subsample=True),
    # Adaptive token sampling (configs/evaluate/vivit_epic_kitchens/_ats.yml): class token, no windows / rel-pos.  The
    # reference's ATS code only runs when batch == heads (blocks.py:163 sums the scores over dim -3, which is the BATCH
    # axis of the (batch, heads, tokens) scores of a 3-D block input, and the per-head index rows are then used per batch
    # entry; ViViT-B meets that by coincidence: 12 batched views, 12 heads), so these cases use batch = heads = 2.
    "tiny_ats_dense": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=2, frames=3, policy=None, block_class="Block",
                           windowed_class=None, has_class_token=True, ats_fraction=0.7, std=0.08, stream="drift", seed=20),
    "tiny_ats_eventful": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=2, frames=5, policy=("fraction", dict(fraction=0.5)),
                              block_class="EventfulBlock", has_class_token=True, ats_fraction=0.7, std=0.08,
                              stream="drift", seed=21),
    # the temporal_ats configuration (configs/evaluate/vivit_epic_kitchens/temporal_ats_200.yml): fp16 attention-value path
    "tiny_ats_cast16": dict(cfg=TINY_GLOBAL, input_size=(4, 4), batch=2, frames=5, policy=("fraction", dict(fraction=0.5)),
                            block_class="EventfulBlock", has_class_token=True, ats_fraction=0.7, matmul_2_cast="float16",
                            std=0.08, stream="drift", seed=22),
}

GATES = ("qkv_gate", "projection_gate", "mlp_gate")


def n_tokens(case):
    h, w = case["input_size"]
    return h * w + int(case.get("has_class_token", False))
