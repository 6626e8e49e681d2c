"""CPU tests: every dataset config of ddpo_b200/config/base.py parses for the experiments it defines, names a registered
reward callback and a prompt function that runs with the configured kwargs; prompt-file loaders; inflect stand-ins."""
import random

import pytest

from ddpo_b200 import training
from ddpo_b200.config import base as C
from ddpo_b200.training import prompts as P
from ddpo_b200.utils import Parser

DATASETS = [n for n, v in vars(C).items() if isinstance(v, dict) and "common" in v]


def test_reference_dataset_names_exist():
    for name in ("compressed_animals", "neg_compressed_animals", "compressed_animals_rwr", "neg_compressed_animals_rwr",
                 "a_animals", "a_animals_rwr", "a_dog_1", "a_dog_2", "llava_bertscore", "llava_counting",
                 "compressed_animals_nocfg", "neg_compressed_animals_nocfg"):
        assert name in DATASETS, name


@pytest.mark.parametrize("dataset", DATASETS)
def test_dataset_config_parses_and_prompts_run(dataset):
    cfg = getattr(C, dataset)
    for experiment in ("pg", "sample", "train"):
        if experiment not in cfg:
            continue
        args = Parser().parse_args(experiment, ["--dataset", dataset])
        assert args.filter_field in training.callback_fns, args.filter_field
        assert args.savepath.startswith(cfg["common"]["logbase"])
        if experiment == "train":        # the fine-tuning driver reads prompts from the stored samples
            continue
        random.seed(0)
        inf, trn, meta = training.make_prompts(args.prompt_fn, 3, getattr(args, "identical_batch", False),
                                               evaluate=False, **args.prompt_kwargs)
        assert len(inf) == len(trn) == len(meta) == 3 and all(isinstance(p, str) and p for p in inf)
        assert all(isinstance(t, (list, tuple)) and t for t in trn)


def test_prompt_functions_follow_the_reference_shapes():
    random.seed(1)
    p, t, m = P.from_file("assets/common_animals.txt", idx=3)
    assert p == P.ANIMALS[3] and t == [p] and m == {}
    p, t, m = P.manual(["a dog", "a cat"])
    assert p in t and t == ["a dog", "a cat"]
    p, t, m = P.vqa_dataset("assets/vqa_debug.txt")
    assert m["prompt"] == p and len(m["questions"]) == len(m["answers"]) == 3
    p, t, m = P.counting("assets/very_simple_animals.txt", (2, 8))
    number, noun = m["answers"]
    assert p == f"{number} {P.plural(noun)}" and m["questions"][0] == f"How many {P.plural(noun)} are there in this image?"
    p, _, _ = P.nouns_activities("assets/common_animals.txt", "assets/activities_v0.txt")
    assert p.split()[0] in ("a", "an") and any(p.endswith(a) for a in P.ACTIVITIES)
    inf, trn, meta = P.make_prompts("imagenet_animals", 4, identical_batch=True)
    assert len(set(inf)) == 1 and trn[0] == [inf[0]]


def test_inflect_stand_ins():
    assert P.indefinite("ant") == "an ant" and P.indefinite("dog") == "a dog"
    assert [P.plural(w) for w in ("cat", "fox", "fly", "mouse", "sheep", "monkey", "wolf")] == \
        ["cats", "foxes", "flies", "mice", "sheep", "monkeys", "wolves"]
    assert P.number_to_words(7) == "seven" and P.number_to_words(42) == "42"
