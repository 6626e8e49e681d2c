"""RWR fine-tuning driver -- mirror of the reference's ``pipeline/finetune.py`` (:46-223): read the filtered samples,
optionally weight them (batch-level softmax or dataset-level ``make_weights``), run ``training.diffusion.train_step``
once per batch, save ``unet_<epoch>.pkl``.

    python -m ddpo_b200.pipeline.finetune --dataset compressed-animals --num_train_epochs 1
"""
import math
import os

import numpy as np
import torch

from .. import datasets, ops, utils
from ..diffusers_patch import DDIMScheduler
from ..training import distributed
from ..training import diffusion as D


class Parser(utils.Parser):
    config = "ddpo_b200.config.base"
    dataset = "compressed_animals"


def main(argv=None, models=None):
    args = Parser().parse_args("train", argv)
    for k, v in (("temperature", 1.0), ("filter_field", None)):
        if not hasattr(args, k):
            setattr(args, k, v)
    from .policy_gradient import set_seed
    set_seed(args.seed)                                                                   # :48
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    # no CUDA device: the kernels refuse CPU tensors (ops._p), so this only serves the CPU dry run of the host logic
    # on the test suite's ops emulator (tests/test_drivers_host_cpu.py)
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    args.modelpath = None if args.iteration == 0 else args.modelpath                      # :51
    if models is None:
        stable_models, stable_params = utils.load_finetuned_stable_diffusion(
            args.modelpath, epoch=args.load_epoch, pretrained_model=args.pretrained_model, dtype=args.dtype,
            cache=args.cache, device=device)
    else:
        pipeline, params = models
        stable_models = utils.serialization.StableModels(pipeline.tokenizer, pipeline.text_encoder, pipeline.vae,
                                                         pipeline.unet)
        stable_params = utils.serialization.StableParams(params.get("vae"), params["unet"])
    tokenizer, text_encoder, vae, unet = stable_models
    print(f"n unet params: {unet.params.numel() / 1e6:.3f}M")

    worker_batch_size = args.train_batch_size * 1                                         # one device per process
    pod_batch_size = worker_batch_size * distributed.world_size()
    train_dataset, train_dataloader = datasets.get_bucket_loader(args.loadpath, tokenizer, batch_size=worker_batch_size,
                                                                 resolution=args.resolution,
                                                                 max_train_samples=args.max_train_samples)
    assert not (args.weighted_batch and args.weighted_dataset), "Cannot weight over both batch and dataset"
    if args.weighted_dataset:
        train_dataset.make_weights(args.filter_field, args.temperature, args.per_prompt_weights)   # :82-85

    tx = D.AdamWConfig(learning_rate=args.learning_rate, b1=args.beta1, b2=args.beta2, eps=args.epsilon,
                       weight_decay=args.weight_decay, max_grad_norm=args.max_grad_norm)           # :89-103
    state = D.TrainState.create(apply_fn=unet, params=unet.params, tx=tx)                         # :104-108
    noise_scheduler = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                    num_train_timesteps=1000, device=device)                      # :110-117 (same a_t table)
    noise_scheduler_state = noise_scheduler.create_state()
    rng = ops.prng_key(args.seed)
    train_rng = ops.threefry_split(rng, 1)[0]                                                     # :128-129
    steps_per_epoch = math.ceil(len(train_dataloader))
    if args.max_train_steps is None:
        args.max_train_steps = args.num_train_epochs * steps_per_epoch
    else:
        args.num_train_epochs = math.ceil(args.max_train_steps / steps_per_epoch)
    static_broadcasted = (noise_scheduler, text_encoder, args.train_cfg, args.guidance_scale)
    print(f"dataset size: {len(train_dataset)} | batch size per device: {args.train_batch_size} | total pod batch "
          f"size: {pod_batch_size} | n epochs: {args.num_train_epochs} | n optimization steps: {args.max_train_steps}")
    global_step, history = 0, []
    for epoch in range(args.num_train_epochs):
        losses = []
        for batch in train_dataloader:
            if args.weighted_batch:
                weights = utils.softmax(np.asarray(batch[args.filter_field]).squeeze(), temperature=args.temperature)
            elif args.weighted_dataset:
                weights = np.asarray(batch["weights"]).squeeze() / pod_batch_size                 # :176-178
            else:
                weights = None
            state, loss, train_rng = D.train_step(state, getattr(text_encoder, "params", None), batch, train_rng,
                                                  noise_scheduler_state, static_broadcasted, weights=weights)
            losses.append(loss)
            global_step += 1
            if global_step >= args.max_train_steps:
                break
        loss_avg = float(torch.stack(losses).mean().item())
        print(f"[ finetune ] epoch {epoch} | average loss {loss_avg:.5f} | steps {global_step}")
        history.append(loss_avg)
        if (epoch + 1) % args.save_freq == 0 or epoch == args.num_train_epochs - 1:               # :214-220
            utils.save_unet(args.savepath, utils.params_tree(state.params, unet.cfg), all_workers=True,
                            epoch=(epoch + 1) // args.save_freq * args.save_freq)
        if global_step >= args.max_train_steps:
            break
    return dict(losses=history, state=state, steps=global_step)


if __name__ == "__main__":
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    main()
