def gca():
    return None
def show():
    pass
