"""Config helpers with the reference's surface (utils/__init__.py:9-48): DictConfig, get_basic_config
(reads configs/basic.yaml relative to the CWD, falling back to the copy shipped in this package),
get_custom_config, get_config, read_lines_from_file."""
import os

import yaml

_PKG_CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'configs', 'basic.yaml')


class DictConfig(object):
    def __init__(self, config_dict):
        self.__dict__.update(config_dict or {})

    def __str__(self):
        return '\n'.join('%s: %s' % kv for kv in self.__dict__.items())

    __repr__ = __str__


def get_custom_config(fname):
    with open(fname, 'r') as stream:
        return DictConfig(yaml.safe_load(stream))


def get_basic_config():
    path = 'configs/basic.yaml'
    return get_custom_config(path if os.path.exists(path) else _PKG_CONFIG)


def get_config(fname):
    config = get_basic_config()
    config.__dict__.update(get_custom_config(fname).__dict__)
    return config


def read_lines_from_file(path, encoding='utf-8'):
    with open(path, 'r', encoding=encoding) as f:
        return [line.strip() for line in f]
