#!/usr/bin/env python
"""CLI with the reference's flags (scripts/tokenize_pdb.py:80-98): tokenizes every .pdb of a directory
into `<stem>_tokens.npy` (uint32, shape (1, n_tokens)) with the CUDA hot path."""
import argparse
import os
import sys
from typing import List, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "protein-structure-tokenizer_b200"))

from pst.config import CODEBOOK_SURNAME, TokenizerConfig, load_config  # noqa: E402
from pst.inference_runner import InferenceRunner  # noqa: E402


def main(pdbs: List[str], token_save_path: str, backend: str, batch_size_per_device: int = 8,
         config_overrides: Optional[List[str]] = None, precision: str = "fp16", random_init: bool = False):
    cfg = load_config(name="vq3d_inference", job_name="tokenize", overrides=config_overrides)
    runner = InferenceRunner()
    local_devices, n_local_device = runner.prepare_devices(backend=backend)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun: one process per GPU; files are sharded over the ranks, rank 0 gathers the tokens over NCCL and writes
        import torch

        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        local_devices = [local_rank]
    if precision != "fp32":
        print(f"NOTE precision={precision}: the edge-level GEMMs use 16-bit operands (fp32 accumulation); >= 99.5 % of the "
              "token ids equal the fp32 reference path's (tests/test_gpu_golden_model.py). Use --precision fp32 for the "
              "all-fp32 path when bit-for-bit comparable ids matter more than speed.", file=sys.stderr)
    tokenize = runner.prepare_tokenize_fn(cfg=cfg, devices=local_devices, precision=precision)
    model_params = runner.load_params(model_dir=cfg.model.weight_paths, local_devices=local_devices, cfg=tokenize.cfg,
                                      allow_random_init=random_init, seed=cfg.random_seed)
    runner.tokenize(random_key=None, quantize=tokenize, model_params=model_params, pdbs=pdbs,
                    token_save_path=token_save_path, data_config=cfg.data.data, num_device=len(local_devices),
                    batch_size_per_device=batch_size_per_device)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch

        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Tokenizer specification !")
    parser.add_argument("--model_downsampling", type=int, choices=[1, 2, 4], default=1)
    parser.add_argument("--codebook_size", type=int, choices=[432, 1728, 4096, 64000], default=4096)
    parser.add_argument("--token_save_path", type=str, required=True)
    parser.add_argument("--pdb_dir", type=str, required=True, help="folder containing the .pdb files to be tokenized")
    parser.add_argument("--backend", type=str, default="gpu", choices=["gpu", "tpu", "cpu"])
    parser.add_argument("--batch_size_per_device", type=int, default=1)
    parser.add_argument("--precision", type=str, default="fp16", choices=["fp32", "fp16", "bf16"],
                        help="operand precision of the edge-level GEMMs. fp16 (default, fastest): >= 99.5 %% of the token ids equal "
                             "the fp32 path's; fp32: all-CUDA-core path whose ids match the reference outside rounding ties")
    parser.add_argument("--random_init", action="store_true", help="use reference-rule random weights if the checkpoint is absent")
    args = parser.parse_args()
    df = args.model_downsampling
    pdbs = [os.path.join(args.pdb_dir, f) for f in os.listdir(args.pdb_dir)]
    model_config = f"gnn/ablation_{CODEBOOK_SURNAME[args.codebook_size]}_df_{df}.yaml"
    overrides = [f"model={model_config}", f"data=ablation_df_{df}.yaml"]
    main(pdbs=pdbs, token_save_path=args.token_save_path, batch_size_per_device=args.batch_size_per_device,
         config_overrides=overrides, backend=args.backend, precision=args.precision, random_init=args.random_init)
