"""Drivers mirroring the reference's ``pipeline/`` scripts: ``policy_gradient`` (DDPO), ``sample`` and ``finetune``
(RWR).  Import the submodule you need; each has a ``main(argv)``."""
