from .bucket import BucketDataset, collate_fn, get_bucket_loader, make_uncond_text  # noqa: F401
