"""Prompt functions -- mirrors the API of the reference's ``ddpo/training/prompts.py``:
``make_prompts(fn_name, batch_size, identical_batch=False, **kwargs) -> (inference_prompts, training_prompts, metadata)``
(:14-34); every prompt function returns ``(inference_prompt, training_prompts, metadata)`` and takes the keyword
arguments the reference's configs pass (``loadpath``, ``prompts``, ``nouns_path`` / ``activities_path``,
``number_range``, ``idx``, ``evaluate``).  Randomness comes from Python's ``random`` (seeded by the drivers, :46).

Data: the reference draws ImageNet class names from ``ddpo/utils/imagenet.py`` and word lists from ``assets/*.txt``;
neither table is part of the accelerated path.  This repo ships its own small lists (``assets/`` at the repo root, and the
``ANIMALS`` fallback below); the ImageNet-based functions sample from ``ANIMALS`` (documented stand-in).  ``inflect`` is
not installed: the three helpers below cover what the prompt functions need from it."""
import os
import random

ANIMALS = ["cat", "dog", "horse", "monkey", "rabbit", "zebra", "spider", "bird", "sheep", "deer", "cow", "goat",
           "lion", "tiger", "bear", "raccoon", "fox", "wolf", "lizard", "beetle", "ant", "butterfly", "fish", "shark",
           "whale", "dolphin", "squirrel", "mouse", "rat", "snake", "turtle", "frog", "chicken", "duck", "goose", "bee",
           "pig", "turkey", "fly", "llama", "camel", "bat", "gorilla", "hedgehog", "kangaroo"]
ACTIVITIES = ["washing the dishes", "riding a bike", "playing chess"]
_NUMBERS = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
            "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen", "twenty"]
_IRREGULAR = {"mouse": "mice", "goose": "geese", "sheep": "sheep", "deer": "deer", "fish": "fish", "wolf": "wolves",
              "fly": "flies", "butterfly": "butterflies"}
_ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "assets")


# ----------------------------------------------------------- small `inflect` stand-ins ----
def indefinite(word):
    """``inflect.engine().a(word)``: 'a dog' / 'an ant'."""
    return f"{'an' if word[:1].lower() in 'aeiou' else 'a'} {word}"


def number_to_words(n):
    return _NUMBERS[n] if 0 <= n < len(_NUMBERS) else str(n)


def plural(noun):
    if noun in _IRREGULAR:
        return _IRREGULAR[noun]
    if noun.endswith(("s", "x", "ch", "sh")):
        return noun + "es"
    if noun.endswith("y") and noun[-2:-1] not in "aeiou":
        return noun[:-1] + "ies"
    return noun + "s"


# ------------------------------------------------------------------- prompt files ----
def _resolve(path):
    """config paths are relative to the repo root in the reference (``assets/common_animals.txt``)"""
    if os.path.exists(path):
        return path
    alt = os.path.join(_ASSETS, os.path.basename(path))
    if os.path.exists(alt):
        return alt
    raise FileNotFoundError(f"prompt file {path} not found (also looked in {_ASSETS})")


def load_lines(loadpath):
    """One prompt per line (reference ``ddpo/utils/serialization.py:511-518``)."""
    with open(_resolve(loadpath), "r") as f:
        return [line.strip() for line in f.readlines()]


def load_general_prompts(path):
    """``PROMPT:`` blocks with SUB / VERB / OBJ question-answer pairs (reference ``serialization.py:484-507``)."""
    dataset = []
    with open(_resolve(path), "r") as f:
        while True:
            line = f.readline()
            if line == "":
                break
            if line == "\n":
                continue
            assert line.startswith("PROMPT: ")
            entry = {"prompt": line[len("PROMPT: "):].strip(), "questions": [], "answers": []}
            for key in ["sub", "verb", "obj"]:
                line = f.readline()
                assert line.startswith(f"{key.upper()} Q: ")
                entry["questions"].append(line[len(f"{key.upper()} Q: "):].strip())
                line = f.readline()
                assert line.startswith(f"{key.upper()} A: ")
                entry["answers"].append(line[len(f"{key.upper()} A: "):].strip())
            dataset.append(entry)
    return dataset


# --------------------------------------------------------------------- general api ----
def batchify(prompt_fn, batch_size, **kwargs):
    inference, training, meta = zip(*[prompt_fn(**kwargs) for _ in range(batch_size)])
    return list(inference), training, meta


def batchify_identical(prompt_fn, batch_size, **kwargs):
    inference, training, meta = prompt_fn(**kwargs)
    return [inference] * batch_size, [training] * batch_size, [meta] * batch_size


def make_prompts(fn_name, batch_size, identical_batch=False, **kwargs):
    prompt_fn = globals()[fn_name]
    return (batchify_identical if identical_batch else batchify)(prompt_fn, batch_size, **kwargs)


# ------------------------------------------------------------- specific experiments ----
def get_random_class(idx=None, low=None, high=None):
    """stand-in for the ImageNet class table: indices wrap into ``ANIMALS``"""
    if idx is not None:
        return ANIMALS[idx % len(ANIMALS)]
    if low is not None and high is not None:
        return ANIMALS[random.randint(low, high) % len(ANIMALS)]
    return random.choice(ANIMALS)


def person_pet(evaluate=False):
    training_prompts = ["a photo of a person with their pet"]
    return random.choice(training_prompts), training_prompts, {}


def consistent_animals(evaluate=False):
    p = "a husky and a shoebill stork on the beach in a single image"
    return p, [p], {}


def n_fingers(evaluate=False):
    n = random.randint(1, 4)
    p = f'a photo of a hand holding up {n} finger{"s" if n > 1 else ""}'
    return p, [p], {}


def imagenet_single(evaluate=False, idx=None):
    p = f"a realistic photo of a {get_random_class(idx=idx)}"
    return p, [p], {}


def imagenet_simple(evaluate=False, idx=None):
    p = f"a {get_random_class(idx=idx)}"
    return p, [p], {}


def imagenet_dogs(evaluate=False, idx=None):
    training_prompts = [f"{get_random_class(idx=idx, low=151, high=268)}"]
    return random.choice(training_prompts), training_prompts, {}


simple_dogs = imagenet_dogs


def imagenet_animals(evaluate=False, idx=None):
    training_prompts = [f"{get_random_class(idx=idx, low=0, high=397)}"]
    return random.choice(training_prompts), training_prompts, {}


def common_animals(evaluate=False, idx=None):
    w = ANIMALS[idx] if idx is not None else random.choice(ANIMALS)
    return w, [w], {}


def animal_debug(evaluate=False, idx=None):
    training_prompts = ["a dog"]
    return random.choice(training_prompts), training_prompts, {}


def from_file(loadpath, evaluate=False, idx=None):
    prompts = load_lines(loadpath)
    p = prompts[idx] if idx is not None else random.choice(prompts)
    return p, [p], {}


def vqa_dataset(loadpath, max_samples=None, evaluate=False):
    entry = random.choice(load_general_prompts(loadpath))
    return entry["prompt"], [entry["prompt"]], entry


def manual(prompts, evaluate=False):
    return random.choice(prompts), prompts, {}


def nouns_activities(nouns_path=None, activities_path=None, evaluate=False):
    nouns = load_lines(nouns_path) if nouns_path else ANIMALS
    activities = load_lines(activities_path) if activities_path else ACTIVITIES
    p = f"{indefinite(random.choice(nouns))} {random.choice(activities)}"
    return p, [p], {}


def counting(nouns_path, number_range, evaluate=False):
    nouns = load_lines(nouns_path)
    number = number_to_words(random.randint(*number_range))
    noun = random.choice(nouns)
    plural_noun = plural(noun)
    p = f"{number} {plural_noun}"
    metadata = {"questions": [f"How many {plural_noun} are there in this image?", "What animal is in this image?"],
                "answers": [number, noun]}
    return p, [p], metadata
