def is_module_wrapper(m):
    return False
