"""`python inference.py --list ... --checkpoint ...` — same command line as the reference's inference.py (its :21-107),
served by tts_arabic_pytorch_b200.inference."""
from tts_arabic_pytorch_b200.inference import main

if __name__ == '__main__':
    main()
