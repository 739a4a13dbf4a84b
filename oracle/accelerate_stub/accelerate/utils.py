def get_balanced_memory(*a, **k):
    return {}


def get_max_memory(*a, **k):
    return {}


def set_module_tensor_to_device(*a, **k):
    return None


def find_tied_parameters(*a, **k):
    return []
