"""RWR sampling driver -- mirror of the reference's ``pipeline/sample.py`` (:1-170): sample a batch with the current
U-Net, decode, score with the ``filter_field`` callback, keep the samples the masker lets through, write shards.

    python -m ddpo_b200.pipeline.sample --dataset compressed-animals --max_steps 4

Deviations (documented in DESIGN.md): shards are local ``.npz`` files (``utils.ShardWriter``; the reference's writer
needs a ``gs://`` bucket); the ``"vae"`` field -- which the reference fills by running the VAE *encoder* over the
decoded image (``callbacks.py:37-57``) -- is filled from the sampled latents themselves (``mean = z / 0.18215``,
``logvar = -30``): they are the exact pre-image of the stored picture, so the lossy decode->encode round trip (and a
VAE encoder on the GPU) is skipped; ``finetune`` then trains on exactly the latent that was sampled."""
import os

import numpy as np
import torch

from .. import datasets, ops, training, utils
from ..training import distributed
from ..training import diffusion as training_diffusion


class Parser(utils.Parser):
    config = "ddpo_b200.config.base"
    dataset = "compressed_animals"


def latents_to_moments(final_latents):
    """[B,4,h,w] sampled latents -> [B,h,w,8] posterior moments (mean | logvar) whose sample is the latent itself."""
    mean = (final_latents.detach().float() / training_diffusion.VAE_SCALING).permute(0, 2, 3, 1)
    return torch.cat([mean, torch.full_like(mean, -30.0)], dim=-1).cpu().numpy()


def main(argv=None, models=None):
    args = Parser().parse_args("sample", argv)
    if args.seed is None:
        args.seed = int(os.environ.get("RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    # no CUDA device: the kernels refuse CPU tensors (ops._p), so this only serves the CPU dry run of the host logic
    # on the test suite's ops emulator (tests/test_drivers_host_cpu.py)
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    rng = ops.prng_key(args.seed)                                                        # :23
    n_devices, process_count = 1, distributed.world_size()
    batch_size = n_devices * args.n_samples_per_device
    print(f"[ sample ] local devices: {n_devices} | pod devices: {n_devices * process_count} | "
          f"worker batch_size: {batch_size} | pod batch size: {batch_size * process_count}")
    loadpath = None if args.iteration == 0 else args.loadpath                            # :37
    if models is None:
        models = utils.load_unet(loadpath, epoch=args.load_epoch, pretrained_model=args.pretrained_model,
                                 cache=args.cache, device=device)
    pipeline, params = models
    pipeline.safety_checker = None
    callback_fns = {args.filter_field: training.callback_fns[args.filter_field]()}       # :47-48
    # the "vae" callback of the reference (:45-48): posterior moments of the DECODED image from the VAE encoder.  A pipeline
    # without an encoder (models passed in by a caller) falls back to the latent's own moments, see latents_to_moments
    vae_encoder = getattr(pipeline, "vae_encoder", None)
    vae_callback = training.callback_fns["vae"](encoder=vae_encoder) if vae_encoder is not None else None
    training_diffusion.patch_scheduler(pipeline)                                         # :55

    writer = utils.ShardWriter(args.savepath, split_size=args.local_size)                # :59-67
    writer.configure("images", encode_fn=utils.encode_jpeg, decode_fn=utils.decode_jpeg)
    writer.configure("inference_prompts")
    writer.configure("training_prompts", encode_fn=utils.encode_generic, decode_fn=utils.decode_generic)
    for key in list(callback_fns) + ["vae"]:
        writer.configure(key)

    def text_encode(ids):
        out = training_diffusion.text_encode(ids, params.get("text_encoder"), pipeline.text_encoder)
        return torch.as_tensor(np.asarray(out, np.float32)).to(device) if not torch.is_tensor(out) else out.to(device)

    uncond_prompt_embeds = text_encode(datasets.make_uncond_text(pipeline.tokenizer, batch_size))   # :71-75
    print(f"[ sample ] embed uncond prompts: {tuple(uncond_prompt_embeds.shape)}")
    masker = utils.make_masker(args.mask_mode, args.mask_param)                          # :79
    avg, timer = utils.StreamingAverage(), utils.Timer()
    print(f"[ sample ] max_samples: {args.max_samples} | max_steps: {args.max_steps} | eval: {args.evaluate}")
    n_steps, n_samples, all_rewards = 0, 0, []
    while True:
        rng, prng_seed = ops.threefry_split(rng, 2)                                      # :92-93
        prng_seeds = ops.threefry_split(prng_seed, n_devices)
        inference_prompts, training_prompts, prompt_metadata = training.make_prompts(
            args.prompt_fn, batch_size, args.identical_batch, evaluate=args.evaluate, **args.prompt_kwargs)
        print(f"[ sample ] prompts: {inference_prompts[:2]}")
        prompt_embeds = text_encode(pipeline.prepare_inputs(inference_prompts))
        final_latents, *_ = pipeline(prompt_embeds, uncond_prompt_embeds, params, prng_seeds[0], args.n_inference_steps,
                                     jit=True, height=args.resolution, width=args.resolution,
                                     guidance_scale=args.guidance_scale, eta=args.eta)    # :108-119
        images = training_diffusion.vae_decode(final_latents, params.get("vae"), pipeline.vae)
        images = images.detach().float().cpu().numpy()
        print(f"[ sample ] {len(images)} samples in {timer():.2f} seconds | eval: {args.evaluate}")
        infos = training.evaluate_callbacks(callback_fns, images, training_prompts, prompt_metadata)
        rewards, metadata = infos[args.filter_field]
        rewards = np.asarray(rewards)
        all_rewards.append(rewards.squeeze())
        avg(rewards.mean().item())
        mask = masker(rewards)                                                           # :139
        batch = {"inference_prompts": inference_prompts, "training_prompts": list(training_prompts), "images": images,
                 "vae": vae_callback(images)[0] if vae_callback is not None else latents_to_moments(final_latents),
                 **{key: np.asarray(rew) for key, (rew, _) in infos.items()}}
        n_added = writer.add_batch(batch, mask=mask)
        n_steps += 1
        tot = torch.tensor([float(n_added)], dtype=torch.float64)
        if distributed.is_distributed():
            tot = tot.to(device)
            torch.distributed.all_reduce(tot)                                            # utils.worker_sum
        n_samples += tot.item()
        print(f"[ sample ] batch {n_steps} / {args.max_steps} | saved: {n_added} | total: {int(n_samples)} / "
              f"{args.max_samples} | average: {avg.avg:.3f} | mask: {masker} | saving: {timer():.2f} seconds\n")
        if args.max_steps is not None and n_steps >= args.max_steps:
            break
        if args.max_samples is not None and n_samples >= args.max_samples:
            break
    writer.close()
    return dict(n_steps=n_steps, n_samples=int(n_samples), rewards=all_rewards, savepath=writer.savepath)


if __name__ == "__main__":
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    main()
