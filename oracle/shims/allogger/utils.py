def report_env(to_stdout=False):
    pass
