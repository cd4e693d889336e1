"""Small helpers with the reference's names.  ref: utils/utilities.py:20-35 (get_cur_time, nan_detection), :213-215."""
import datetime
import pickle

import numpy as np


def get_cur_time():
    return datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S')


def nan_detection(val_name, val):
    """Aborts on a NaN cost, like the reference (utils/utilities.py:33-35)."""
    if np.isnan(val):
        assert False, '[ERROR] {} is nan.'.format(val_name)


def pickle_dump(filename, obj):
    with open(filename, 'wb') as fp:
        pickle.dump(obj, fp, 2)
