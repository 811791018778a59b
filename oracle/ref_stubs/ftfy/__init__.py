"""Import shim (test infrastructure only): the reference's tokenizer imports ftfy.fix_text;
the CLIPSelf hot path never tokenizes, so identity is sufficient."""


def fix_text(text):
    return text
