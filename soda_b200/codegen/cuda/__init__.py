def add_arguments(parser): pass
def print_code(stencil, args): pass
