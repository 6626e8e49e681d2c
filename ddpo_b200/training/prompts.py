"""Prompt functions -- mirrors the API of ``ddpo/training/prompts.py:14-34``:
``make_prompts(fn_name, batch_size, identical_batch=False, **kwargs) -> (inference_prompts,
training_prompts, metadata)``; each prompt fn returns ``(inference_prompt, training_prompts, metadata)``.
The word lists are this repo's own small synthetic stand-ins (the reference's asset files / ImageNet table
are data, not part of the accelerated path)."""
import random

ANIMALS = ["cat", "dog", "horse", "monkey", "rabbit", "zebra", "spider", "bird", "sheep", "deer", "cow", "goat",
           "lion", "tiger", "bear", "raccoon", "fox", "wolf", "lizard", "beetle", "ant", "butterfly", "fish", "shark",
           "whale", "dolphin", "squirrel", "mouse", "rat", "snake", "turtle", "frog", "chicken", "duck", "goose", "bee",
           "pig", "turkey", "fly", "llama", "camel", "bat", "gorilla", "hedgehog", "kangaroo"]
ACTIVITIES = ["washing the dishes", "riding a bike", "playing chess"]


def batchify(prompt_fn, batch_size, **kwargs):
    inference, training, meta = zip(*[prompt_fn(**kwargs) for _ in range(batch_size)])
    return list(inference), training, meta


def batchify_identical(prompt_fn, batch_size, **kwargs):
    inference, training, meta = prompt_fn(**kwargs)
    return [inference] * batch_size, [training] * batch_size, [meta] * batch_size


def make_prompts(fn_name, batch_size, identical_batch=False, **kwargs):
    prompt_fn = globals()[fn_name]
    return (batchify_identical if identical_batch else batchify)(prompt_fn, batch_size, **kwargs)


def _pick(words, evaluate):
    w = random.choice(words)
    return w, [w], {}


def common_animals(evaluate=False):
    return _pick(ANIMALS, evaluate)


def imagenet_animals(evaluate=False):
    return _pick(ANIMALS, evaluate)


def animal_debug(evaluate=False):
    return "a dog", ["a dog"], {}


def nouns_activities(nouns_path=None, activities_path=None, evaluate=False):
    a, act = random.choice(ANIMALS), random.choice(ACTIVITIES)
    article = "an" if a[0] in "aeiou" else "a"
    p = f"{article} {a} {act}"
    return p, [p], {}
