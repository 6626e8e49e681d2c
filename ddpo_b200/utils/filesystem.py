"""Local-filesystem subset of the reference's ``ddpo/utils/filesystem.py`` (``join_and_create`` :100-104,
``is_remote`` :65-66, ``save_json``/``read_json`` :90-97, ``unpickle`` :60-62, ``ls`` :25-36, ``exists`` :39-44).
The reference writes to Google Cloud Storage (``gs://`` paths, ``gcsfs``): there is no network on the GPU box, so a
``gs://bucket/x`` path is mapped to the local mirror ``logs/bucket/x`` -- the same rule the reference applies for its
local copies (``pipeline/policy_gradient.py:90``: ``"logs/" + savepath.replace("gs://", "")``)."""
import json
import os
import pickle


def is_remote(path):
    return isinstance(path, str) and path.startswith("gs://")


def localize(path, cache="logs"):
    return os.path.join(cache, path.replace("gs://", "")) if is_remote(path) else path


def join_and_create(*args):
    path = os.path.join(*args)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    return path


def exists(path):
    return os.path.exists(localize(path))


def ls(path):
    return sorted(os.listdir(localize(path)))


def save_json(path, x):
    with open(join_and_create(localize(path)), "w") as f:
        json.dump(x, f, indent=4)


def read_json(path):
    with open(localize(path)) as f:
        return json.load(f)


def unpickle(path):
    with open(localize(path), "rb") as f:
        return pickle.load(f)
