"""``python -m spine_vision_b200 dataset {classification,localization} [flags]`` -- the two dataset sub-commands of the
reference's ``spine-vision`` CLI (``spine_vision/cli/__init__.py:30-56, 110-126``) over the GPU path.  The flags are generated
from the same configuration fields the reference hands to tyro (kebab-case names, ``--flag`` / ``--no-flag`` booleans,
tuples as several values: ``--crop-size 128 128 --crop-delta-mm 50 20 30 30``).  Training and the Phenikaa OCR pre-processing
stay with the reference (not on this path)."""

from __future__ import annotations

import argparse
import logging
import sys
import typing
from pathlib import Path


def _add_fields(parser: argparse.ArgumentParser, cfg_cls) -> None:
    for name, field in cfg_cls.model_fields.items():
        flag = "--" + name.replace("_", "-")
        ann, default = field.annotation, field.default
        origin, args = typing.get_origin(ann), typing.get_args(ann)
        if ann is bool:
            parser.add_argument(flag, action=argparse.BooleanOptionalAction, default=default)
        elif origin is tuple:
            parser.add_argument(flag, nargs=len(args), type=args[0], default=default, metavar=name[0].upper())
        elif origin is typing.Literal:
            parser.add_argument(flag, choices=list(args), default=default)
        elif origin in (typing.Union, getattr(__import__("types"), "UnionType", ())) and type(None) in args:
            inner = [a for a in args if a is not type(None)][0]
            parser.add_argument(flag, type=inner, default=default)
        else:
            parser.add_argument(flag, type=ann if ann in (int, float, str, Path) else str, default=default)


def main(argv=None) -> int:
    from .dataset import ClassificationDatasetConfig, create_classification_dataset
    from .localization_dataset import LocalizationDatasetConfig, create_localization_dataset

    top = argparse.ArgumentParser(prog="python -m spine_vision_b200", description=__doc__.split("\n\n")[0])
    sub = top.add_subparsers(dest="command", required=True)
    ds = sub.add_parser("dataset", help="dataset creation on the GPU path").add_subparsers(dest="kind", required=True)
    commands = {"classification": (ClassificationDatasetConfig, create_classification_dataset, "Create classification dataset (Phenikaa + SPIDER)"),
                "localization": (LocalizationDatasetConfig, create_localization_dataset, "Create localization dataset")}
    for kind, (cfg_cls, _, text) in commands.items():
        _add_fields(ds.add_parser(kind, help=text, description=text), cfg_cls)
    ns = vars(top.parse_args(argv))
    cfg_cls, run, _ = commands[ns.pop("kind")]
    ns.pop("command")
    for k, v in list(ns.items()):
        if isinstance(v, list):
            ns[k] = tuple(v)
    config = cfg_cls(**ns)
    logging.basicConfig(level=logging.DEBUG if config.verbose else logging.INFO, format="%(levelname)s %(name)s: %(message)s")
    result = run(config)
    print(result.summary)
    return 0


if __name__ == "__main__":
    sys.exit(main())
