class AlignDevicesHook:
    def __init__(self, *a, **k):
        pass


def add_hook_to_module(module, hook, *a, **k):
    return module


def remove_hook_from_submodules(module, *a, **k):
    return None
