"""Write-request tags of the operator protocol (reference: mobula/const.py:1-5)."""


class req:
    null = 'null'
    write = 'write'
    inplace = 'inplace'
    add = 'add'
