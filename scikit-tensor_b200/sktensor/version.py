__version__ = '0.1+b200.1'
