"""Boundary plumbing of the reference's util.py that the train path needs: prepare_dirs_and_logger (util.py:17-47)
and save_config (util.py:52-59).  Plotting / image helpers (util.py:64-280) are visualisation and out of scope."""
import json
import logging
import os
from datetime import datetime


def prepare_dirs_and_logger(config):
    # (the reference chdir()s to its own source directory, util.py:19; paths here stay relative to the caller's cwd)
    formatter = logging.Formatter("%(asctime)s:%(levelname)s::%(message)s")
    logger = logging.getLogger()
    for hdlr in list(logger.handlers):
        logger.removeHandler(hdlr)
    handler = logging.StreamHandler()
    handler.setFormatter(formatter)
    logger.addHandler(handler)

    config.data_path = os.path.join(config.data_dir, config.dataset)     # util.py:33
    if config.load_path:                                                 # util.py:37-38
        config.model_dir = config.load_path
    elif not hasattr(config, 'model_dir'):                               # util.py:40-44 (a preset model_dir is kept)
        model_name = "{}/{}_{}_{}".format(config.dataset, datetime.now().strftime("%m%d_%H%M%S"), config.arch,
                                          config.tag)
        config.model_dir = os.path.join(config.log_dir, model_name)
    os.makedirs(config.model_dir, exist_ok=True)


def save_config(config):
    param_path = os.path.join(config.model_dir, "params.json")
    print("[*] MODEL dir: %s" % config.model_dir)
    print("[*] PARAM path: %s" % param_path)
    with open(param_path, 'w') as fp:
        json.dump(config.__dict__, fp, indent=4, sort_keys=True)
